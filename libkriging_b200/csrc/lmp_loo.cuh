// lmp_loo.cuh -- LMP / LOO specific stages and predict (f1).
// Included inside engine.cu's anonymous namespace, after struct Engine.
//
// Both objectives need the same rank-p corrected inverse
//     Bm = R^-1 - U U^T ,   U = R^-1 F LX^-T ,   LX LX^T = F^T R^-1 F
// * LMP (src/lib/Kriging.cpp:488-648): Bm is `R^-1 - Rinv_X_Xt_Rinv_X_inv_Xt_Rinv`; the per-k
//   `Wb_k` products of compute_lmp_theta_ans (src/lib/KrigingImpl.cpp:887-921) collapse to
//   ans_k = -1/2 sum_ij G_k,ij Bm_ij + Qo^T G_k Qo / (2 sigma2): the pair reduction of cov.cuh
//   with weight Bm (U passed to the kernel, Bm never materialised).
// * LOO (src/lib/Kriging.cpp:353-468): the reference's B = Linv^T Linv - A^T A with
//   A = Qstar^T Linv equals Bm (Qstar Qstar^T = Fstar (Fstar^T Fstar)^-1 Fstar^T).  With
//   s = 1/diag(Bm), By = x, e = s.x, the d per-k products diagABA(B, gradR_k) (2 n^3 each)
//   collapse to ONE symmetric product M = Bm diag(s^3 x^2) Bm (n^3) followed by the same pair
//   reduction:   dloo/dtheta_k = (2/n) sum_{l != j} G_k,lj (M_lj - v_l x_j),  v = Bm (e.s).

__global__ void __launch_bounds__(256)
cross_partial_kernel(const double* __restrict__ Am, long long lda, int qa, const double* __restrict__ Bmat,
                     long long ldb, int qb, int n, double* __restrict__ partial) {
  // partial[chunk][i * qb + j] = sum_{rows in chunk} A[r, i] B[r, j]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * GRAM_CHUNK;
  const int r1 = min(n, r0 + GRAM_CHUNK);
  for (int pr = warp; pr < qa * qb; pr += 8) {
    const int i = pr / qb, j = pr % qb;
    const double* za = Am + (long long)i * lda;
    const double* zb = Bmat + (long long)j * ldb;
    double s = 0.0;
    for (int r = r0 + lane; r < r1; r += 32) s += za[r] * zb[r];
    s = warp_sum(s);
    if (lane == 0) partial[(long long)blockIdx.x * qa * qb + pr] = s;
  }
}

// U[r, j] = sum_i Z[r, i] T[i, j]  (T: p x p column-major), rows < n; also q[r] = yv[r] - sum_j U[r, j] w[j]
__global__ void __launch_bounds__(256)
right_mult_small_kernel(const double* __restrict__ Z, long long ldz, const double* __restrict__ T, int p, int n,
                        double* __restrict__ U, long long ldu, const double* __restrict__ yv,
                        const double* __restrict__ w, double* __restrict__ q) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double acc = yv ? yv[r] : 0.0;
  for (int j = 0; j < p; ++j) {
    double s = 0.0;
    for (int i = 0; i <= j; ++i) s += Z[(long long)i * ldz + r] * T[(long long)j * p + i];
    U[(long long)j * ldu + r] = s;
    if (yv) acc -= s * w[j];
  }
  if (q) q[r] = acc;
}

// LOO per-row quantities.  bdiag = V_ii - |u_i|^2 ; s = 1/bdiag ; e = s x ; c = s^3 x^2 ; es = e s.
// partial[cta] = sum e^2 over the CTA's rows.
__global__ void __launch_bounds__(256)
loo_rows_kernel(const double* __restrict__ V, long long ld, const double* __restrict__ U, long long ldu, int p,
                const double* __restrict__ xvec, int n, int N, double* __restrict__ s2loo, double* __restrict__ err,
                double* __restrict__ sqrtc, double* __restrict__ es, double* __restrict__ partial) {
  __shared__ double sh[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e2 = 0.0;
  if (i < N) {
    if (i < n) {
      double b = V[(long long)i * ld + i];
      for (int q = 0; q < p; ++q) {
        const double u = U[(long long)q * ldu + i];
        b -= u * u;
      }
      const double s = 1.0 / b, x = xvec[i];
      const double e = s * x;
      s2loo[i] = s;
      err[i] = e;
      sqrtc[i] = sqrt(fabs(s * s * s)) * fabs(x);
      es[i] = e * s;
      e2 = e * e;
    } else {
      s2loo[i] = 0.0;
      err[i] = 0.0;
      sqrtc[i] = 0.0;
      es[i] = 0.0;
    }
  }
  e2 = warp_sum(e2);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

// Q[i, j] = sqrtc[i] * (Vsym[i, j] - u_i . u_j) for the full N x N square (V holds the lower block triangle).
// 64 x 64 tiles; the transposed read of the lower tile goes through shared memory.
__global__ void __launch_bounds__(256)
scaled_bm_kernel(const double* __restrict__ V, long long ld, const double* __restrict__ U, long long ldu, int p,
                 const double* __restrict__ sqrtc, int n, int N, double* __restrict__ Q) {
  __shared__ double tile[64][65];
  const int tpr = N / 64;
  const int ti = blockIdx.x % tpr, tj = blockIdx.x / tpr;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // ty in 0..3
  if (ti >= tj) {
    for (int c = ty; c < 64; c += 4) tile[c][tx] = V[(long long)(tj * 64 + c) * ld + ti * 64 + tx];  // tile[c][r]
  } else {
    for (int c = ty; c < 64; c += 4) tile[tx][c] = V[(long long)(ti * 64 + c) * ld + tj * 64 + tx];  // transpose
  }
  __syncthreads();
  const int i = ti * 64 + tx;
  const double sc = sqrtc[i];
  for (int c = ty; c < 64; c += 4) {
    const int j = tj * 64 + c;
    double v = tile[c][tx];
    if (i < n && j < n)
      for (int q = 0; q < p; ++q) v -= U[(long long)q * ldu + i] * U[(long long)q * ldu + j];
    Q[(long long)j * ld + i] = sc * v;
  }
}

// out = Vsym * w - U (U^T w) for a lower-block-triangle-stored symmetric V: one warp per row pair sweep.
// Stage 1: per column j (one warp): t_j = sum_{i >= j} V[i, j] w_i  (lower incl. diag) and scatter of the strict
// lower part is avoided by computing the row form separately: out_i = sum_{j <= i} V[i,j] w_j + sum_{j > i} V[j,i] w_j.
__global__ void __launch_bounds__(256)
symv_lower_kernel(const double* __restrict__ V, long long ld, const double* __restrict__ w, int n,
                  double* __restrict__ out) {
  // one warp per output row i.  The row part (j <= i) is a strided read; acceptable for an O(n^2) side stage
  // executed once per LOO gradient.
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = 0.0;
  for (int j = lane; j < i; j += 32) s += V[(long long)j * ld + i] * w[j];
  const double* col = V + (long long)i * ld;
  for (int r = i + lane; r < n; r += 32) s += col[r] * w[r];
  s = warp_sum(s);
  if (lane == 0) out[i] = s;
}

// v[i] -= sum_q U[i, q] t[q]
__global__ void rank_p_correct_kernel(double* __restrict__ v, const double* __restrict__ U, long long ldu, int p,
                                      const double* __restrict__ t, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = v[i];
  for (int q = 0; q < p; ++q) s -= U[(long long)q * ldu + i] * t[q];
  v[i] = s;
}

// predict: rectangular correlation block R_on[i, j] = rho(x_i - xn_j) * factor (1.0 when the points coincide,
// src/lib/KrigingImpl.cpp:193-221), rows padded to N with zeros.
template <int KERNEL>
__global__ void __launch_bounds__(256)
cov_rect_kernel(const double* __restrict__ X, int n, int d, const double* __restrict__ Xn, int m,
                const __grid_constant__ KernelParams kp, double factor, double* __restrict__ Ron, long long ldr, int N) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * m) return;
  const int j = (int)(idx / N), i = (int)(idx % N);
  double v = 0.0;
  if (i < n) {
    double es = 0.0, pr = 1.0;
    bool zero = true;
    for (int k = 0; k < d; ++k) {
      const double dx = X[(long long)k * n + i] - Xn[(long long)k * m + j];
      zero = zero && (fabs(dx) <= 2.220446049250313e-16);
      corr_accum<KERNEL>(dx * (kp.inv_theta[k] * corr_scale<KERNEL>()), es, pr);
    }
    v = zero ? 1.0 : corr_finish<KERNEL>(es, pr) * factor;
  }
  Ron[(long long)j * ldr + i] = v;
}

// per new point j: out[j][0] = sum_r S[r,j]^2 ; out[j][1] = sum_r S[r,j] z[r] ; out[j][2+q] = sum_r S[r,j] Fstar[r,q]
// for q < p, and out[j][2+p] = sum_r S[r,j] ystar[r]  (ystar = column p of [Fstar | ystar]; used when beta is fixed)
__global__ void __launch_bounds__(256)
predict_dots_kernel(const double* __restrict__ S, long long lds, const double* __restrict__ z,
                    const double* __restrict__ Fstar, long long ldf, int p, int n, double* __restrict__ out) {
  __shared__ double sh[8];
  const int j = blockIdx.x;
  const double* col = S + (long long)j * lds;
  for (int q = -2; q <= p; ++q) {
    double s = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const double a = col[r];
      const double b = (q == -2) ? a : (q == -1 ? z[r] : Fstar[(long long)q * ldf + r]);
      s += a * b;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += sh[w];
      out[(long long)j * (p + 3) + q + 2] = t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// Lower Cholesky of a small p x p matrix with the reference's safe_chol_lower semantics
// (src/lib/LinearAlgebra.cpp:43-98).  The condition number is the exact 1-norm one (explicit inverse).
static void host_safe_chol_lower(std::vector<double>& G, int p, double num_nugget, int max_inc, double min_rcond,
                                 bool rcond_check, std::vector<double>& Linv) {
  std::vector<double> X = G, L((size_t)p * p), Li((size_t)p * p);
  int inc = 0;
  while (true) {
    bool ok = true;
    std::fill(L.begin(), L.end(), 0.0);
    for (int j = 0; j < p && ok; ++j) {
      double s = X[(size_t)j * p + j];
      for (int k = 0; k < j; ++k) s -= L[(size_t)k * p + j] * L[(size_t)k * p + j];
      if (!(s > 0.0)) {
        ok = false;
        break;
      }
      const double ljj = std::sqrt(s);
      L[(size_t)j * p + j] = ljj;
      for (int i = j + 1; i < p; ++i) {
        double t = X[(size_t)j * p + i];
        for (int k = 0; k < j; ++k) t -= L[(size_t)k * p + i] * L[(size_t)k * p + j];
        L[(size_t)j * p + i] = t / ljj;
      }
    }
    bool wrong = false;
    if (ok) {
      // inverse by forward substitution, column by column
      std::fill(Li.begin(), Li.end(), 0.0);
      for (int c = 0; c < p; ++c)
        for (int i = c; i < p; ++i) {
          double t = (i == c) ? 1.0 : 0.0;
          for (int k = c; k < i; ++k) t -= L[(size_t)k * p + i] * Li[(size_t)c * p + k];
          Li[(size_t)c * p + i] = t / L[(size_t)i * p + i];
        }
      if (rcond_check) {
        double n1 = 0.0, n2 = 0.0;
        for (int c = 0; c < p; ++c) {
          double a = 0.0, b = 0.0;
          for (int i = c; i < p; ++i) {
            a += std::fabs(L[(size_t)c * p + i]);
            b += std::fabs(Li[(size_t)c * p + i]);
          }
          n1 = std::max(n1, a);
          n2 = std::max(n2, b);
        }
        const double rc = 1.0 / (n1 * n2);
        wrong = !(rc * rc >= min_rcond);
      }
    }
    if (!ok || wrong) {
      if (inc > max_inc) throw LkError{"[ERROR] Exceed max numerical nugget added to force chol matrix (F' R^-1 F)"};
      if (num_nugget <= 0.0) throw LkError{"[ERROR] Cannot add numerical nugget which is not strictly positive"};
      for (int j = 0; j < p; ++j) X[(size_t)j * p + j] += num_nugget * std::pow(10.0, inc);
      ++inc;
      continue;
    }
    break;
  }
  G = L;
  Linv = Li;
}

// Common LMP / LOO preparation.  On return: Uv = U (N x p), Qv = Qo = R^-1 y - U U^T y, host scalars filled.
void Engine::lmp_loo_prepare(int objective, int& pdim) {
  (void)objective;
  pdim = p;
  const int q = p + 1;
  // [Rinv_X | yt_Rinv] = L^T \ [Fstar | ystar]
  dev_copy(Zv, Bv, (long long)N * q);
  solve_bwd(Zv, q);
  // H = [F | y]^T [Rinv_X | yt_Rinv]   (q x q)
  const int chunks = (n + GRAM_CHUNK - 1) / GRAM_CHUNK;
  launches += 3;
  cross_partial_kernel<<<chunks, 256, 0, s_main>>>(dF, n, p, Zv, N, q, n, dpartial);
  sum_partials_kernel<<<(p * q + 63) / 64, 64, 0, s_main>>>(dpartial, chunks, p * q, p * q, dsmall);
  cross_partial_kernel<<<chunks, 256, 0, s_main>>>(dy, n, 1, Zv, N, q, n, dpartial + (size_t)chunks * p * q);
  launches += 1;
  sum_partials_kernel<<<(q + 63) / 64, 64, 0, s_main>>>(dpartial + (size_t)chunks * p * q, chunks, q, q, dsmall + p * q);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaMemcpyAsync(hpin, dsmall, (size_t)q * q * 8, cudaMemcpyDeviceToHost, s_main));
  CUDA_CHECK(cudaStreamSynchronize(s_main));
  // hpin[i*q + j] (i < p): F_i . Z_j ;  hpin[p*q + j]: y . Z_j
  std::vector<double> G((size_t)p * p), LXinv;
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j) G[(size_t)j * p + i] = hpin[i * q + j];  // Xt_Rinv_X = F^T Rinv_X (column-major)
  const double yRy = hpin[p * q + p];
  std::vector<double> Fty(p);
  for (int j = 0; j < p; ++j) Fty[j] = hpin[p * q + j];  // y^T Rinv_X
  host_safe_chol_lower(G, p, num_nugget, max_inc, min_rcond, rcond_check, LXinv);
  double sumlog = 0.0;
  for (int j = 0; j < p; ++j) sumlog += std::log(G[(size_t)j * p + j]);
  // w = LX^-1 (Rinv_X^T y) = U^T y
  std::vector<double> w(p, 0.0);
  double wtw = 0.0;
  for (int i = 0; i < p; ++i) {
    double s = 0.0;
    for (int k = 0; k <= i; ++k) s += LXinv[(size_t)k * p + i] * Fty[k];
    w[i] = s;
    wtw += s * s;
  }
  h_sum_log_diagLX = sumlog;
  h_S2 = yRy - wtw;
  // T = LX^-T (upper), U = Rinv_X T ; Qo = yt_Rinv - U w
  std::vector<double> T((size_t)p * p, 0.0);
  for (int c = 0; c < p; ++c)
    for (int r = 0; r <= c; ++r) T[(size_t)c * p + r] = LXinv[(size_t)r * p + c];
  memcpy(hpin, T.data(), (size_t)p * p * 8);
  memcpy(hpin + (size_t)p * p, w.data(), (size_t)p * 8);
  CUDA_CHECK(cudaMemcpyAsync(dsmall, hpin, ((size_t)p * p + p) * 8, cudaMemcpyHostToDevice, s_main));
  dev_zero(Uv, (long long)N * p);
  dev_zero(Qv, N);
  launches += 1;
  right_mult_small_kernel<<<(n + 255) / 256, 256, 0, s_main>>>(Zv, N, dsmall, p, n, Uv, N, Zv + (size_t)N * p,
                                                               dsmall + (size_t)p * p, Qv);
  CUDA_CHECK(cudaGetLastError());
  // Uv / dsmall are consumed asynchronously; hpin may be reused only after the next sync (eval's final sync
  // happens before any further host write to hpin).
  CUDA_CHECK(cudaStreamSynchronize(s_main));
}

// LOO value and (optionally) gradient.  Requires V = R^-1 (lower tiles), Uv, Xv (= By).
void Engine::loo_finish(int want_grad) {
  const int g = (N + 255) / 256;
  launches += 2;
  loo_rows_kernel<<<g, 256, 0, s_main>>>(V, ld, Uv, N, p, Xv, n, N, dS2loo, dErr, dSqrtC, dEs, dpartial);
  sum_partials_kernel<<<1, 32, 0, s_main>>>(dpartial, g, 1, 1, dscal + SC_LOO);
  CUDA_CHECK(cudaGetLastError());
  have_loo = true;
  if (!want_grad) return;
  if (!Q1) {
    Q1 = dalloc<double>((size_t)N * N);
    Q2 = dalloc<double>((size_t)N * N);
    mapQ1 = maps_of(Q1, false);
  }
  // v = Bm (e.s)
  launches += 3;
  symv_lower_kernel<<<(n * 32 + 255) / 256, 256, 0, s_main>>>(V, ld, dEs, n, Tv);
  dev_zero(Tv + n, N - n);
  cross_partial_kernel<<<(n + GRAM_CHUNK - 1) / GRAM_CHUNK, 256, 0, s_main>>>(Uv, N, p, dEs, N, 1, n, dpartial);
  sum_partials_kernel<<<(p + 63) / 64, 64, 0, s_main>>>(dpartial, (n + GRAM_CHUNK - 1) / GRAM_CHUNK, p, p, dsmall);
  launches += 1;
  rank_p_correct_kernel<<<(n + 255) / 256, 256, 0, s_main>>>(Tv, Uv, N, p, dsmall, n);
  // Q1 = diag(sqrt c) Bm (full square) ;  Q2 = Q1^T Q1 = Bm diag(c) Bm (lower tiles)
  launches += 1;
  scaled_bm_kernel<<<(N / 64) * (N / 64), 256, 0, s_main>>>(V, ld, Uv, N, p, dSqrtC, n, N, Q1);
  CUDA_CHECK(cudaGetLastError());
  {
    if (!loo_table) {
      // every tile walks the whole k range: rounds of the persistent grid are blocks of lauum_band rows x
      // ~(2 SMs / lauum_band) columns that start together and stay together (see build_plans, LAUUM order 3)
      std::vector<TileDesc> lt;
      if (lauum_order == 3) {
        lt = tables::lower_tiles_in_bands(nb, N, lauum_band, false);
      } else {
        lt = tables::lower_tiles_by_row(nb, N, false);
        if (l2_order) order_for_l2(lt);
      }
      loo_tiles = (int)lt.size();
      loo_table = dalloc<TileDesc>(lt.size());
      CUDA_CHECK(cudaMemcpy(loo_table, lt.data(), lt.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    }
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.C = Q2;
    a.ldc = ld;
    a.sched = SCHED_TABLE;
    a.epilogue = EPI_SET;
    a.table = loo_table;
    a.ntiles = loo_tiles;
    gemm(2, mapQ1, Q1, mapQ1, Q1, a, s_main, true);
  }
  // pair reduction with weight M = Q2 and the rank-1 term (v, x)
  grad_reduce(1.0, 0, SC_GRAD, Q2, Tv, Xv);
}

// predict mean / variance factor (src/lib/KrigingImpl.cpp:145-243) from the model of the last evaluation.
void Engine::predict(int m, const double* Xn, const double* Fn, const double* beta, double r_on_factor,
                     double* mean_out, double* var_out) {
  CUDA_CHECK(cudaSetDevice(device));
  if (!have_model) throw LkError{"lkgpu_predict: no evaluation has been run on this handle"};
  if (m < 1) throw LkError{"lkgpu_predict: need m >= 1"};
  SweepGate gate(*this);
  for (int k = 0; k < d; ++k) kp.inv_theta[k] = 1.0 / last_theta[k];
  std::vector<double> hRstar((size_t)p * p);
  if (p > 0) CUDA_CHECK(cudaMemcpy(hRstar.data(), dRstar, (size_t)p * p * 8, cudaMemcpyDeviceToHost));
  const int chunk_max = 1024;
  double* dXn = dalloc<double>((size_t)std::min(m, chunk_max) * d);
  double* dS = dalloc<double>((size_t)N * std::min(m, chunk_max));
  double* dout = dalloc<double>((size_t)std::min(m, chunk_max) * (p + 3));
  std::vector<double> hx((size_t)chunk_max * d), ho((size_t)chunk_max * (p + 3));
  try {
    for (int j0 = 0; j0 < m; j0 += chunk_max) {
      const int mc = std::min(chunk_max, m - j0);
      for (int k = 0; k < d; ++k)
        for (int j = 0; j < mc; ++j) hx[(size_t)k * mc + j] = Xn[(size_t)k * m + j0 + j];
      CUDA_CHECK(cudaMemcpyAsync(dXn, hx.data(), (size_t)mc * d * 8, cudaMemcpyHostToDevice, s_main));
      const long long tot = (long long)N * mc;
      ++launches;
#define LAUNCH_RECT(K) \
  cov_rect_kernel<K><<<(unsigned)((tot + 255) / 256), 256, 0, s_main>>>(dX, n, d, dXn, mc, kp, r_on_factor, dS, N, N)
      switch (kernel) {
        case 0: LAUNCH_RECT(0); break;
        case 1: LAUNCH_RECT(1); break;
        case 2: LAUNCH_RECT(2); break;
        default: LAUNCH_RECT(3); break;
      }
#undef LAUNCH_RECT
      CUDA_CHECK(cudaGetLastError());
      solve_fwd(dS, mc);  // Rstar_on = L \ R_on
      ++launches;
      predict_dots_kernel<<<mc, 256, 0, s_main>>>(dS, N, Ev, Bv, N, p, n, dout);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaMemcpyAsync(ho.data(), dout, (size_t)mc * (p + 3) * 8, cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaStreamSynchronize(s_main));
      std::vector<double> e(p);
      for (int j = 0; j < mc; ++j) {
        const double* o = ho.data() + (size_t)j * (p + 3);
        // estimated beta: z = Estar; fixed beta (lkgpu_set_fixed_beta): z = ystar - Fstar beta (Kriging.cpp:2168-2172)
        double mean = has_fixed_beta ? o[2 + p] : o[1];
        if (has_fixed_beta)
          for (int q = 0; q < p; ++q) mean -= o[2 + q] * fixed_beta[q];
        for (int q = 0; q < p; ++q) mean += Fn[(size_t)q * m + j0 + j] * beta[q];
        mean_out[j0 + j] = mean;
        if (var_out) {
          // Ecirc = (Fn - Rstar_on' M) circ^-1 : row-vector solve with the upper factor
          double ss = 0.0;
          for (int q = 0; q < p; ++q) {
            double t = Fn[(size_t)q * m + j0 + j] - o[2 + q];
            for (int k = 0; k < q; ++k) t -= e[k] * hRstar[(size_t)q * p + k];
            e[q] = t / hRstar[(size_t)q * p + q];
            ss += e[q] * e[q];
          }
          double v = 1.0 - o[0] + ss;
          if (!(v >= 0.0)) v = 0.0;
          var_out[j0 + j] = v;
        }
      }
    }
  } catch (...) {
    cudaFree(dXn);
    cudaFree(dS);
    cudaFree(dout);
    throw;
  }
  cudaFree(dXn);
  cudaFree(dS);
  cudaFree(dout);
}
