/* Plain-C restatement of the integer/byte part of the hot path: clip normalisation + voxelization.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as nm_oracle.py): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; nothing under neural_marionette_b200/ does.  It exists as a second,
 * numpy-free statement of the arithmetic the CUDA kernel `nm_normalize_voxelize` must reproduce bit for bit,
 * and is itself pinned by tests/test_oracle_golden.py against the fixtures written from the reference's own
 * outputs (tests/golden/voxelize_hashes.json, voxelize_obj.npz).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: every float32 / float64 operation rounds once, as numpy's do)
 *
 * Reference followed (file:line under /root/reference):
 *   utils/dataset_utils.py:9-19   episodic_normalization
 *   utils/dataset_utils.py:21-31  voxelize (is_binarized=True branch)
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* utils/dataset_utils.py:11-15 for a float32 clip `seq` of n points (T*N rows of xyz).
 * numpy >= 2 keeps the whole expression in float32 (python scalars are weak) up to the final
 * `+ np.array([x_trans, 0, z_trans])`, which promotes to float64.  out: n*3 doubles. */
int nmo_episodic_normalization_f32(const float* seq, size_t n, double scale, double x_trans, double z_trans, double* out) {
  if (n == 0) return 1;                                   /* np.amax of an empty array raises */
  float lo[3], hi[3];
  for (int d = 0; d < 3; d++) lo[d] = hi[d] = seq[d];
  for (size_t i = 1; i < n; i++)
    for (int d = 0; d < 3; d++) {
      const float v = seq[3 * i + d];
      if (v < lo[d]) lo[d] = v;
      if (v > hi[d]) hi[d] = v;
    }
  float blen = hi[0] - lo[0];                             /* :14  (bmax - bmin).max() */
  for (int d = 1; d < 3; d++) { const float e = hi[d] - lo[d]; if (e > blen) blen = e; }
  const float den = blen + (float)1e-5;                   /* :15  blen + 1e-5 (float32) */
  const float s = (float)scale;
  const double trans[3] = {x_trans, 0.0, z_trans};
  for (size_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) {
      float v = seq[3 * i + d] - lo[d];
      v = v * s;
      v = v / den;
      v = v * 2.0f;
      v = v - 1.0f;
      out[3 * i + d] = (double)v + trans[d];
    }
  return 0;
}

/* utils/dataset_utils.py:24-29 for n float64 points with `stride` doubles per row (>= 3; extra columns such as
 * normals are ignored, :27).  grid: G*G*G floats, index [ix][iy][iz], zero-filled here (:26).
 * Returns 0, or 2 if an index falls outside [0, G) (numpy would wrap a negative index / raise on a large one;
 * the product asserts instead). */
int nmo_voxelize_f64(const double* pts, size_t n, size_t stride, int G, float* grid) {
  memset(grid, 0, sizeof(float) * (size_t)G * G * G);
  const double step = (1.0 - (-1.0)) / (double)G;         /* :25  (bbox[3:] - bbox[:3]) / output_shape */
  const double den = step + 1e-5;                         /* :28 */
  int bad = 0;
  for (size_t i = 0; i < n; i++) {
    int32_t c[3];
    for (int d = 0; d < 3; d++) c[d] = (int32_t)((pts[stride * i + d] - (-1.0)) / den);   /* truncation toward zero */
    if (c[0] < 0 || c[1] < 0 || c[2] < 0 || c[0] >= G || c[1] >= G || c[2] >= G) { bad = 2; continue; }
    grid[((size_t)c[0] * G + c[1]) * G + c[2]] = 1.0f;    /* :29  idempotent store */
  }
  return bad;
}

/* Callers' loop (vis_generation.py:15-23): raw float32 clip (T, N, 3) -> (T, G, G, G) float32 occupancy.
 * scratch: T*N*3 doubles. */
int nmo_normalize_voxelize_clip(const float* raw, int T, size_t N, int G, double scale, double x_trans, double z_trans,
                                double* scratch, float* grids) {
  int rc = nmo_episodic_normalization_f32(raw, (size_t)T * N, scale, x_trans, z_trans, scratch);
  if (rc) return rc;
  for (int t = 0; t < T; t++) {
    const int r = nmo_voxelize_f64(scratch + (size_t)t * N * 3, N, 3, G, grids + (size_t)t * G * G * G);
    if (r) rc = r;
  }
  return rc;
}

/* utils/eval_utils.py:43-46 for ONE frame: symmetric chamfer distance between the occupied voxels of gt (non-zero) and
 * of recon (>= 0.5, the reference binarises first, :37-38).  Coordinates idx / ((G-1)/2) - 1 and squared distances in
 * float32 as torch computes them; the two means are accumulated in double (torch sums float32 pairwise: agreement to
 * ~1e-7 relative, the test allows 1e-6).  Returns 1 when either set is empty (torch raises there).  O(n_gt * n_recon). */
int nmo_voxel_chamfer_frame(const float* gt, const float* recon, int G, double* out) {
  const size_t S = (size_t)G * G * G;
  size_t na = 0, nb = 0;
  for (size_t i = 0; i < S; i++) { na += gt[i] != 0.0f; nb += recon[i] >= 0.5f; }
  if (na == 0 || nb == 0) return 1;
  const float half = (float)((G - 1) / 2.0);
  double sum_a = 0.0, sum_b = 0.0;
  for (int dir = 0; dir < 2; dir++) {
    const float* q = dir ? recon : gt;
    const float* t = dir ? gt : recon;
    double acc = 0.0;
    for (size_t i = 0; i < S; i++) {
      if (!(dir ? q[i] >= 0.5f : q[i] != 0.0f)) continue;
      const float qx = (float)(i / ((size_t)G * G)) / half - 1.0f, qy = (float)((i / G) % G) / half - 1.0f,
                  qz = (float)(i % G) / half - 1.0f;
      float best = 3.4e38f;
      for (size_t j = 0; j < S; j++) {
        if (!(dir ? t[j] != 0.0f : t[j] >= 0.5f)) continue;
        const float dx = qx - ((float)(j / ((size_t)G * G)) / half - 1.0f), dy = qy - ((float)((j / G) % G) / half - 1.0f),
                    dz = qz - ((float)(j % G) / half - 1.0f);
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best) best = d;
      }
      acc += (double)best;
    }
    if (dir) sum_b = acc / (double)nb; else sum_a = acc / (double)na;
  }
  *out = sum_a + sum_b;
  return 0;
}

/* utils/eval_utils.py:65-81 for ONE frame: keypoints (K, 4) with intensity < threshold count as (1e4, 1e4, 1e4); idx_out[k']
 * = the detected keypoint nearest to gt joint k' (first minimum, float32 distances summed x, y, z in that order). */
void nmo_semantic_nearest_frame(const float* kypt, const float* gt, int K, int Kgt, float threshold, int* idx_out) {
  for (int g = 0; g < Kgt; g++) {
    float best = 0.0f;
    int arg = 0;
    for (int k = 0; k < K; k++) {
      const int invalid = kypt[4 * k + 3] < threshold;
      const float x = invalid ? 1e4f : kypt[4 * k], y = invalid ? 1e4f : kypt[4 * k + 1], z = invalid ? 1e4f : kypt[4 * k + 2];
      const float dx = gt[3 * g] - x, dy = gt[3 * g + 1] - y, dz = gt[3 * g + 2] - z;
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (k == 0 || d < best) { best = d; arg = k; }
    }
    idx_out[g] = arg;
  }
}
