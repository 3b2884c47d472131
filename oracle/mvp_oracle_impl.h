/*
 * TEST INFRASTRUCTURE ONLY — body of the CPU oracle, included once per scalar type by
 * mvp_oracle.c (T = float / double).  See mvp_oracle.c for the contract.
 *
 * Required macros: T (scalar type), SFX (symbol suffix), FMA(a,b,c) (fused multiply-add in T),
 * SQ3(dx,dy,dz) / SQ2(dx,dy) (squared norm in the reference's contraction order).
 *
 * Arithmetic note (reference: mvpnet/ops/setup.py:22 builds with plain `nvcc -O2`, i.e. the
 * default -fmad=true): every `dist = 0; dist += diff * diff` loop in the reference kernels is
 * contracted by nvcc.  What it contracts TO was read off the SASS of the reference sources built
 * unmodified for sm_100a with nvcc 12.9 (oracle/build_ref.py, `cuobjdump -sass oracle/_ref/*.so`)
 * and is confirmed bit-for-bit by tests/test_gpu_vs_reference_kernels.py:
 *     float :  FFMA(dx,dx,0) ; FFMA(dy,dy,.) ; FFMA(dz,dz,.)   ==  fma(dz,dz, fma(dy,dy, dx*dx))
 *     double:  DMUL(dy,dy)   ; DFMA(dx,dx,.) ; DFMA(dz,dz,.)   ==  fma(dz,dz, fma(dx,dx, dy*dy))
 * (in double the `0.0 +` is folded away and, of the two products then being added, the FIRST one
 * is fused while the second is rounded).  2-D points (FPS only): float fma(dy,dy, dx*dx), double
 * fma(dx,dx, dy*dy).  The oracle spells these chains out so that near-threshold / near-tie cases
 * resolve exactly as on the GPU.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* squared distance with the reference's contraction order; diff = key - query */
static inline T FN(sqdist3)(const T *a, const T *q) {
  T dx = a[0] - q[0], dy = a[1] - q[1], dz = a[2] - q[2];
  return SQ3(dx, dy, dz);
}

/* ------------------------------------------------------------------------------------------
 * Farthest point sampling.  Follows mvpnet/ops/cuda/fps_kernel.cu:60-135 (kernel) and
 * :144-180 (host: block size rule :21-24, temp initialised to -1 :160, index zeros :158).
 * The block's threads are simulated one by one, then the shared-memory tree reduction
 * (:117-129) is replayed literally, so the reference tie rule
 *   (max dist, then smallest j mod BLOCK, then smallest j)
 * is reproduced by construction rather than by a derived formula.
 * ------------------------------------------------------------------------------------------ */
int FN(mvpo_fps)(const T *points, int64_t B, int64_t N, int64_t D, int64_t M, int64_t *index) {
  if (D != 2 && D != 3) return 1;
  if (M <= 0 || N < M) return 2;
  const int BS = mvpo_ref_block_size(N, 16);
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < B; ++b) {
    const T *pts = points + b * N * D;
    int64_t *out = index + b * M;
    T *temp = (T *)malloc(sizeof(T) * (size_t)N);
    T *sd = (T *)malloc(sizeof(T) * (size_t)BS);
    int *si = (int *)malloc(sizeof(int) * (size_t)BS);
    if (!temp || !sd || !si) { err = 3; free(temp); free(sd); free(si); continue; }
    for (int64_t j = 0; j < N; ++j) temp[j] = (T)-1.0;
    int cur = 0;
    out[0] = 0;
    for (int64_t i = 1; i < M; ++i) {
      T c[3] = {0, 0, 0};
      for (int d = 0; d < D; ++d) c[d] = pts[(int64_t)cur * D + d];
      for (int t = 0; t < BS; ++t) {
        T max_dist = (T)0.0;
        int max_idx = cur;
        for (int64_t j = t; j < N; j += BS) {
          T dist;
          if (D == 3) dist = SQ3(pts[j * 3] - c[0], pts[j * 3 + 1] - c[1], pts[j * 3 + 2] - c[2]);
          else dist = SQ2(pts[j * 2] - c[0], pts[j * 2 + 1] - c[1]);
          T last = temp[j];
          if (last > dist || last < (T)0.0) temp[j] = dist; else dist = last;
          if (dist > max_dist) { max_dist = dist; max_idx = (int)j; }
        }
        sd[t] = max_dist;
        si[t] = max_idx;
      }
      for (int off = BS / 2; off > 0; off /= 2)
        for (int t = 0; t < off; ++t)
          if (sd[t] < sd[t + off]) { sd[t] = sd[t + off]; si[t] = si[t + off]; }
      cur = si[0];
      out[i] = cur;
    }
    free(temp); free(sd); free(si);
  }
  return err;
}

/* ------------------------------------------------------------------------------------------
 * Ball query (optionally with distances).  Follows mvpnet/ops/cuda/ball_query_kernel.cu:58-135
 * (+ host fill with -1 at :164) and ball_query_distance_kernel.cu:59-139 (+ :169-171).
 *  - radius arrives as a C `float` and is squared in T (ball_query_kernel.cu:73,150)
 *  - keys scanned in index order, strict `<`, first K hits kept
 *  - the tail is padded with the FIRST hit (index only; distances stay -1)
 *  - a query with no hit keeps its row of -1
 * ------------------------------------------------------------------------------------------ */
int FN(mvpo_ball_query)(const T *query, const T *key, int64_t B, int64_t N1, int64_t N2,
                        float radius, int64_t K, int64_t *index, T *distance) {
  if (K <= 0) return 1;
  const T r = (T)radius;
  const T r2 = r * r;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t i = 0; i < N1; ++i) {
      const T *q = query + (b * N1 + i) * 3;
      const T *kb = key + b * N2 * 3;
      int64_t *row = index + (b * N1 + i) * K;
      T *drow = distance ? distance + (b * N1 + i) * K : NULL;
      for (int64_t k = 0; k < K; ++k) { row[k] = -1; if (drow) drow[k] = (T)-1.0; }
      int64_t cnt = 0;
      for (int64_t j = 0; j < N2 && cnt < K; ++j) {
        T d = FN(sqdist3)(kb + j * 3, q);
        if (d < r2) { row[cnt] = j; if (drow) drow[cnt] = d; ++cnt; }
      }
      if (cnt < K) { int64_t pad = row[0]; for (int64_t k = cnt; k < K; ++k) row[k] = pad; }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * 3-NN with squared distances.  Follows mvpnet/ops/cuda/knn_distance_kernel.cu:35-124 including
 * the literal initialisation `min_dist[K] = {1e40}` / `min_idx[K] = {-1}` (:67-68: only element 0
 * is set; the rest are zero) and the strict-`<` sorted insertion (:96-107).
 * ------------------------------------------------------------------------------------------ */
int FN(mvpo_knn3)(const T *query, const T *key, int64_t B, int64_t N1, int64_t N2,
                  int64_t *index, T *distance) {
  if (N2 < 3) return 1;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t i = 0; i < N1; ++i) {
      const T *q = query + (b * N1 + i) * 3;
      const T *kb = key + b * N2 * 3;
      T md[3] = {(T)1e40, (T)0, (T)0};
      int mi[3] = {-1, 0, 0};
      for (int64_t j = 0; j < N2; ++j) {
        T d = FN(sqdist3)(kb + j * 3, q);
        for (int k = 0; k < 3; ++k) {
          if (d < md[k]) {
            for (int l = 2; l > k; --l) { md[l] = md[l - 1]; mi[l] = mi[l - 1]; }
            md[k] = d; mi[k] = (int)j;
            break;
          }
        }
      }
      for (int k = 0; k < 3; ++k) {
        index[(b * N1 + i) * 3 + k] = mi[k];
        distance[(b * N1 + i) * 3 + k] = md[k];
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * group_points forward / backward.  Follows mvpnet/ops/cuda/group_points_kernel.cu:25-47
 * (forward == gather along the last axis) and :50-89 (backward == scatter-add; the reference
 * uses atomicAdd so its summation order is unspecified — the oracle sums in (n, k) order and
 * tests compare floating-point sums with a tolerance).
 * Returns 4 on an out-of-range index (the reference device-asserts, :85).
 * ------------------------------------------------------------------------------------------ */
int FN(mvpo_group_points_fwd)(const T *in, const int64_t *index, int64_t B, int64_t C, int64_t N1,
                              int64_t N2, int64_t K, T *out) {
  int err = 0;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      const T *src = in + (b * C + c) * N1;
      const int64_t *idx = index + b * N2 * K;
      T *dst = out + (b * C + c) * N2 * K;
      for (int64_t e = 0; e < N2 * K; ++e) {
        int64_t j = idx[e];
        if (j < 0 || j >= N1) { err = 4; dst[e] = 0; } else dst[e] = src[j];
      }
    }
  return err;
}

int FN(mvpo_group_points_bwd)(const T *gout, const int64_t *index, int64_t B, int64_t C,
                              int64_t N1, int64_t N2, int64_t K, T *gin) {
  int err = 0;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      T *dst = gin + (b * C + c) * N1;
      const int64_t *idx = index + b * N2 * K;
      const T *src = gout + (b * C + c) * N2 * K;
      for (int64_t j = 0; j < N1; ++j) dst[j] = 0;
      for (int64_t e = 0; e < N2 * K; ++e) {
        int64_t j = idx[e];
        if (j < 0 || j >= N1) err = 4; else dst[j] += src[e];
      }
    }
  return err;
}

/* ------------------------------------------------------------------------------------------
 * feature_interpolate forward / backward (K == 3).  Follows
 * mvpnet/ops/cuda/interpolate_kernel.cu:25-68 (out = sum_k in[idx_k] * w_k, accumulated
 * k = 0,1,2 with nvcc's fma contraction) and :131-174 (atomic scatter of grad*w).
 * ------------------------------------------------------------------------------------------ */
int FN(mvpo_interpolate_fwd)(const T *in, const int64_t *index, const T *weight, int64_t B,
                             int64_t C, int64_t M, int64_t N, T *out) {
  int err = 0;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      const T *src = in + (b * C + c) * M;
      T *dst = out + (b * C + c) * N;
      for (int64_t n = 0; n < N; ++n) {
        const int64_t *idx = index + (b * N + n) * 3;
        const T *w = weight + (b * N + n) * 3;
        T acc = (T)0.0;
        for (int k = 0; k < 3; ++k) {
          int64_t j = idx[k];
          if (j < 0 || j >= M) { err = 4; continue; }
          acc = FMA(src[j], w[k], acc);
        }
        dst[n] = acc;
      }
    }
  return err;
}

int FN(mvpo_interpolate_bwd)(const T *gout, const int64_t *index, const T *weight, int64_t B,
                             int64_t C, int64_t M, int64_t N, T *gin) {
  int err = 0;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      T *dst = gin + (b * C + c) * M;
      const T *src = gout + (b * C + c) * N;
      for (int64_t j = 0; j < M; ++j) dst[j] = 0;
      for (int64_t n = 0; n < N; ++n) {
        const int64_t *idx = index + (b * N + n) * 3;
        const T *w = weight + (b * N + n) * 3;
        for (int k = 0; k < 3; ++k) {
          int64_t j = idx[k];
          if (j < 0 || j >= M) { err = 4; continue; }
          dst[j] += src[n] * w[k];
        }
      }
    }
  return err;
}

#undef FN
#undef CAT
#undef CAT_
