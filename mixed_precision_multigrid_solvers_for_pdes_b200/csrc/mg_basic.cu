// libmgb200: basic per-operator kernels (strict reference arithmetic, no alignment requirements).
//
// Each kernel mirrors one method of the reference's operator protocol (see include/mgb200.h for
// the file:line map).  They use true divisions and non-contracted adds/muls, so fp64 AND fp32
// results equal the reference's NumPy results bit for bit on every input; the temporally
// blocked / fused kernels in mg_vcycle.cu are validated against these on large grids.
//
// Thread mapping: threadIdx.x walks j (the contiguous index), rows are spread over
// blockIdx.y/threadIdx.y, so every warp reads and writes contiguous row segments.
#include <stdio.h>
#include <atomic>
#include "mg_common.cuh"
#include "mg_lexgs.cuh"

namespace mg {

static thread_local char g_err[256];
static std::atomic<long long> g_launches{0};

int check_launch(const char* what, int launches) {
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return MG_ERR_LAUNCH;
  }
  return MG_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

constexpr int BX = 128;  // threads along j
constexpr int BY = 2;    // rows per block

static inline dim3 grid2d(int nx, int ny) { return dim3((ny + BX - 1) / BX, (nx + BY - 1) / BY); }

// ------------------------------------------------------------------------------------------
// A u  and  r = f - A u
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BX* BY) apply_kernel(const T* __restrict__ u, T* __restrict__ out, int nx,
                                                        int ny, int64_t ldu, int64_t ldo, StencilScalars<T> s) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  T v = (T)0;
  if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
    const T* p = u + (int64_t)i * ldu + j;
    v = apply_strict<T>(s, p[0], p[ldu], p[-ldu], p[1], p[-1]);
  }
  out[(int64_t)i * ldo + j] = v;
}

// TI = storage type of u and f, TO = storage type of r, TC = compute type (wider of the two).
template <typename TI, typename TO, typename TC>
__global__ void __launch_bounds__(BX* BY)
    residual_kernel(const TI* __restrict__ u, const TI* __restrict__ f, TO* __restrict__ r, int nx, int ny,
                    int64_t ldu, int64_t ldf, int64_t ldr, StencilScalars<TC> s) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  TC v = (TC)f[(int64_t)i * ldf + j];
  if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
    const TI* p = u + (int64_t)i * ldu + j;
    v = Strict<TC>::sub(v, apply_strict<TC>(s, (TC)p[0], (TC)p[ldu], (TC)p[-ldu], (TC)p[1], (TC)p[-1]));
  }
  r[(int64_t)i * ldr + j] = (TO)v;
}

// ------------------------------------------------------------------------------------------
// Smoothers
// ------------------------------------------------------------------------------------------
// One colour of red-black GS, in place.  Thread (i, jj) owns the interior point
// j = 2*jj + 1 + ((i + 1 + colour) & 1)  so no thread idles on the other colour.
template <typename T>
__global__ void __launch_bounds__(BX* BY) rbgs_colour_kernel(T* __restrict__ u, const T* __restrict__ f, int nx,
                                                              int ny, int64_t ldu, int64_t ldf, int colour,
                                                              StencilScalars<T> s) {
  const int jj = blockIdx.x * BX + threadIdx.x, i = 1 + blockIdx.y * BY + threadIdx.y;
  if (i >= nx - 1) return;
  const int j = 2 * jj + 1 + ((i + 1 + colour) & 1);
  if (j >= ny - 1) return;
  T* p = u + (int64_t)i * ldu + j;
  p[0] = relax_strict<T>(s, p[0], p[ldu], p[-ldu], p[1], p[-1], f[(int64_t)i * ldf + j]);
}

template <typename T>
__global__ void __launch_bounds__(BX* BY) jacobi_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                         const T* __restrict__ f, int nx, int ny, int64_t ldu,
                                                         int64_t ldf, StencilScalars<T> s) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  const T* p = src + (int64_t)i * ldu + j;
  T v = p[0];
  if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1)
    v = relax_strict<T>(s, v, p[ldu], p[-ldu], p[1], p[-1], f[(int64_t)i * ldf + j]);
  dst[(int64_t)i * ldu + j] = v;
}

// Lexicographic GS along anti-diagonals d = i + j (one block; see header).
template <typename T>
__device__ __forceinline__ void lexgs_sweep_block(T* u, const T* f, int nx, int ny, int64_t ldu, int64_t ldf,
                                                  const StencilScalars<T>& s, bool backward = false) {
  const int dlast = nx + ny - 4;
  for (int dd = 2; dd <= dlast; ++dd) {
    const int d = backward ? (dlast + 2 - dd) : dd;
    const int ilo = max(1, d - (ny - 2)), ihi = min(nx - 2, d - 1);
    for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
      const int j = d - i;
      T* p = u + (int64_t)i * ldu + j;
      p[0] = relax_strict<T>(s, p[0], p[ldu], p[-ldu], p[1], p[-1], f[(int64_t)i * ldf + j]);
    }
    __syncthreads();
  }
}

// Coarsest-level solve, base.py:258-285 in one launch.  When SMEM, u and f live in shared memory
// (dense ld = ny) for the whole iteration.
template <typename T, bool SMEM>
__global__ void coarse_solve_kernel(T* gu, const T* gf, int nx, int ny, int64_t ldu, int64_t ldf, double hxhy,
                                    double tol, int max_it, double* info, StencilScalars<T> s) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  T* u = gu;
  const T* f = gf;
  int64_t lu = ldu, lf = ldf;
  if (SMEM) {
    T* su = reinterpret_cast<T*>(smem_raw);
    T* sf = su + nx * ny;
    for (int k = threadIdx.x; k < nx * ny; k += blockDim.x) {
      const int i = k / ny, j = k - i * ny;
      su[k] = gu[(int64_t)i * ldu + j];
      sf[k] = gf[(int64_t)i * ldf + j];
    }
    __syncthreads();
    u = su; f = sf; lu = ny; lf = ny;
  }
  int it = 0;
  double norm = 0.0;
  for (it = 1; it <= max_it; ++it) {
    lexgs_sweep_block<T>(u, f, nx, ny, lu, lf, s);
    // r = f - A u (boundary r = f), sum of squares over all points; the squares are formed in T
    // like NumPy's field**2, the accumulation is fp64
    double acc = 0.0;
    for (int k = threadIdx.x; k < nx * ny; k += blockDim.x) {
      const int i = k / ny, j = k - i * ny;
      T v = f[(int64_t)i * lf + j];
      if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
        const T* p = u + (int64_t)i * lu + j;
        v = Strict<T>::sub(v, apply_strict<T>(s, p[0], p[lu], p[-lu], p[1], p[-1]));
      }
      acc += (double)Strict<T>::mul(v, v);
    }
    acc = block_reduce(acc, red);
    norm = sqrt(hxhy * acc);
    if (norm < tol) break;  // uniform across the block
  }
  if (it > max_it) it = max_it;
  if (SMEM) {
    __syncthreads();
    for (int k = threadIdx.x; k < nx * ny; k += blockDim.x) {
      const int i = k / ny, j = k - i * ny;
      if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) gu[(int64_t)i * ldu + j] = u[k];
    }
  }
  if (info != nullptr && threadIdx.x == 0) {
    info[0] = (double)it;
    info[1] = norm;
  }
}

// ------------------------------------------------------------------------------------------
// Transfers
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(BX* BY) restrict_kernel(const TI* __restrict__ fine, TO* __restrict__ coarse,
                                                           int nxc, int nyc, int64_t ldf, int64_t ldc, int method) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nxc || j >= nyc) return;
  using A = Strict<TI>;
  const TI* p = fine + (int64_t)(2 * i) * ldf + 2 * j;
  TI v = p[0];
  const bool interior = (i > 0 && i < nxc - 1 && j > 0 && j < nyc - 1);
  if (interior && method != MG_RESTRICT_INJECTION) {
    const TI edges = A::add(A::add(A::add(p[-ldf], p[ldf]), p[-1]), p[1]);
    if (method == MG_RESTRICT_FULL_WEIGHTING) {
      const TI corners = A::add(A::add(A::add(p[-ldf - 1], p[-ldf + 1]), p[ldf - 1]), p[ldf + 1]);
      v = A::add(A::add(A::mul((TI)(1.0 / 16.0), corners), A::mul((TI)(1.0 / 8.0), edges)),
                 A::mul((TI)(1.0 / 4.0), p[0]));
    } else {
      v = A::add(A::mul((TI)(1.0 / 8.0), edges), A::mul((TI)(1.0 / 2.0), p[0]));
    }
  }
  coarse[(int64_t)i * ldc + j] = (TO)v;
}

// One thread per FINE point; values are formed in the fine dtype TO from coarse values cast to TO
// (transfer.py:236-265).
template <typename TI, typename TO>
__global__ void __launch_bounds__(BX* BY) prolong_kernel(const TI* __restrict__ coarse, TO* __restrict__ fine,
                                                          int nxf, int nyf, int64_t ldc, int64_t ldf, int method,
                                                          int add) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nxf || j >= nyf) return;
  using A = Strict<TO>;
  const int ic = i >> 1, jc = j >> 1;
  const TI* c = coarse + (int64_t)ic * ldc + jc;
  TO v = (TO)0;
  const bool oi = i & 1, oj = j & 1;
  if (!oi && !oj) {
    v = (TO)c[0];
  } else if (method == MG_PROLONG_BILINEAR) {
    if (oi && !oj) {
      if (j < nyf - 1) v = A::mul((TO)0.5, A::add((TO)c[0], (TO)c[ldc]));
    } else if (!oi && oj) {
      if (i < nxf - 1) v = A::mul((TO)0.5, A::add((TO)c[0], (TO)c[1]));
    } else {
      v = A::mul((TO)0.25, A::add(A::add(A::add((TO)c[0], (TO)c[1]), (TO)c[ldc]), (TO)c[ldc + 1]));
    }
  }
  TO* o = fine + (int64_t)i * ldf + j;
  o[0] = add ? A::add(o[0], v) : v;
}

// ------------------------------------------------------------------------------------------
// Reductions and element-wise helpers
// ------------------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 1024;  // fixed partial count => fixed summation tree
constexpr int RED_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(RED_THREADS) sumsq_partial_kernel(const T* __restrict__ x, int nx, int ny,
                                                                    int64_t ld, double* __restrict__ partial) {
  __shared__ double red[32];
  double acc = 0.0;
  // block b owns rows b, b + gridDim.x, ...
  for (int i = blockIdx.x; i < nx; i += gridDim.x) {
    const T* row = x + (int64_t)i * ld;
    for (int j = threadIdx.x; j < ny; j += RED_THREADS) {
      const T v = row[j];
      acc += (double)Strict<T>::mul(v, v);
    }
  }
  acc = block_reduce(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

template <bool MAX>
__global__ void __launch_bounds__(RED_THREADS) final_reduce_kernel(const double* __restrict__ partial, int n,
                                                                   double* __restrict__ out) {
  __shared__ double red[32];
  double acc = MAX ? -1.0 : 0.0;
  for (int k = threadIdx.x; k < n; k += RED_THREADS) acc = MAX ? fmax(acc, partial[k]) : acc + partial[k];
  acc = block_reduce<MAX>(acc, red);
  if (threadIdx.x == 0) out[0] = acc;
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(BX* BY) cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int nx,
                                                       int ny, int64_t lds, int64_t ldd) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  dst[(int64_t)i * ldd + j] = (TD)src[(int64_t)i * lds + j];
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(BX* BY) axpy_kernel(TY alpha, const TX* __restrict__ x, TY* __restrict__ y,
                                                       int nx, int ny, int64_t ldx, int64_t ldy) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  TY* o = y + (int64_t)i * ldy + j;
  o[0] = Strict<TY>::add(o[0], Strict<TY>::mul(alpha, (TY)x[(int64_t)i * ldx + j]));
}

__device__ __forceinline__ double sinsin_value(int i, int j, int nx, int ny, double x0, double x1, double y0,
                                               double y1, double amp, double kx, double ky) {
  // np.linspace: start + k*step, end point exact (core/grid.py:48-49)
  const double sx = (x1 - x0) / (double)(nx - 1), sy = (y1 - y0) / (double)(ny - 1);
  const double x = (i == nx - 1) ? x1 : __dadd_rn(__dmul_rn((double)i, sx), x0);
  const double y = (j == ny - 1) ? y1 : __dadd_rn(__dmul_rn((double)j, sy), y0);
  return __dmul_rn(__dmul_rn(amp, sin(__dmul_rn(kx * M_PI, x))), sin(__dmul_rn(ky * M_PI, y)));
}

template <typename T>
__global__ void __launch_bounds__(BX* BY) fill_sinsin_kernel(T* __restrict__ f, int nx, int ny, int64_t ld,
                                                              double x0, double x1, double y0, double y1,
                                                              double amp, double kx, double ky) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  f[(int64_t)i * ld + j] = (T)sinsin_value(i, j, nx, ny, x0, x1, y0, y1, amp, kx, ky);
}

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
    maxerr_partial_kernel(const T* __restrict__ u, int nx, int ny, int64_t ld, double x0, double x1, double y0,
                          double y1, double amp, double kx, double ky, double* __restrict__ partial) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = blockIdx.x; i < nx; i += gridDim.x)
    for (int j = threadIdx.x; j < ny; j += RED_THREADS)
      acc = fmax(acc, fabs((double)u[(int64_t)i * ld + j] - sinsin_value(i, j, nx, ny, x0, x1, y0, y1, amp, kx, ky)));
  acc = block_reduce<true>(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// dst (mapped pinned host memory, addressed through its device alias) = src (device): a store over PCIe by the SMs
__global__ void store_doubles_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

// zero the first / last column and optionally the first / last row: thread t handles row t and column t
template <typename T>
__global__ void zero_ring_kernel(T* __restrict__ x, int nx, int ny, int64_t ld, int first_row, int last_row) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nx) {
    x[(int64_t)t * ld] = (T)0;
    x[(int64_t)t * ld + ny - 1] = (T)0;
  }
  if (t < ny) {
    if (first_row) x[t] = (T)0;
    if (last_row) x[(int64_t)(nx - 1) * ld + t] = (T)0;
  }
}

// Right-hand side of one theta-method heat step (docs/methodology.md:710 divided by theta*dt), ONE pass:
//   rhs = lam * ( u + c_lap * L_h u + c_f1 * f1 + c_f0 * f0 ),   L_h = lap_h (a == null) or div(a grad .) (nodal a),
// L_h u = 0 on the first / last local row and column (a slab's outer ghost rows lose one row of validity, like
// mg_apply_laplacian), the boundary ring zeroed where the slab touches the physical boundary, and the sum of rhs^2 over
// rows [row_lo, row_hi) (the relative stopping test of the step) as one partial per block.  Strict arithmetic.
__global__ void __launch_bounds__(RED_THREADS)
    heat_rhs_kernel(const double* __restrict__ u, const double* __restrict__ f1, const double* __restrict__ f0,
                    const double* __restrict__ a, double* __restrict__ rhs, int nx, int ny, int64_t ldu, int64_t ldf1,
                    int64_t ldf0, int64_t lda, int64_t ldr, double ihx2, double ihy2, double c_lap, double c_f1, double c_f0,
                    double lam, int zero_first, int zero_last, int row_lo, int row_hi, double* __restrict__ partial) {
  using A = Strict<double>;
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = blockIdx.x; i < nx; i += gridDim.x) {
    const double* ur = u + (int64_t)i * ldu;
    for (int j = threadIdx.x; j < ny; j += RED_THREADS) {
      double t = ur[j];
      if (c_lap != 0.0 && i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
        double lap;
        if (a != nullptr) {
          const double* ar = a + (int64_t)i * lda + j;
          const double ae = A::mul(0.5, A::add(ar[lda], ar[0])), aw = A::mul(0.5, A::add(ar[-lda], ar[0]));
          const double an = A::mul(0.5, A::add(ar[1], ar[0])), as = A::mul(0.5, A::add(ar[-1], ar[0]));
          const double x = A::mul(A::add(A::mul(ae, A::sub(ur[j + ldu], t)), A::mul(aw, A::sub(ur[j - ldu], t))), ihx2);
          const double y = A::mul(A::add(A::mul(an, A::sub(ur[j + 1], t)), A::mul(as, A::sub(ur[j - 1], t))), ihy2);
          lap = A::add(x, y);
        } else {
          const double x = A::mul(A::add(ur[j + ldu], ur[j - ldu]), ihx2), y = A::mul(A::add(ur[j + 1], ur[j - 1]), ihy2);
          lap = A::sub(A::add(x, y), A::mul(t, A::add(A::mul(2.0, ihx2), A::mul(2.0, ihy2))));
        }
        t = A::add(t, A::mul(c_lap, lap));
      }
      if (f1 != nullptr) t = A::add(t, A::mul(c_f1, f1[(int64_t)i * ldf1 + j]));
      if (f0 != nullptr) t = A::add(t, A::mul(c_f0, f0[(int64_t)i * ldf0 + j]));
      t = A::mul(lam, t);
      if (j == 0 || j == ny - 1 || (zero_first && i == 0) || (zero_last && i == nx - 1)) t = 0.0;
      rhs[(int64_t)i * ldr + j] = t;
      if (i >= row_lo && i < row_hi) acc += A::mul(t, t);
    }
  }
  acc = block_reduce(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

void reduce_partials_sum(const double* partials, int n, double* out, cudaStream_t st) {
  final_reduce_kernel<false><<<1, RED_THREADS, 0, st>>>(partials, n, out);
}

}  // namespace mg

// ==============================================================================================
// C ABI
// ==============================================================================================
using namespace mg;

static inline bool valid_dtype(int d) { return d == MG_F32 || d == MG_F64; }

template <typename T>
static int launch_lexgs(T* u, const T* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                        double omega, int sweeps, int mode, cudaStream_t st) {
  const int nwarps = (nx - 2 + 31) / 32, blocks = nwarps;
  const auto sc = make_scalars<T>(hx, hy, omega, 1.0);
  int launches = 0;
  for (int k = 0; k < sweeps; ++k)
    for (int dir = 0; dir < 2; ++dir) {
      if ((dir == 0 && mode == MG_LEXGS_BACKWARD) || (dir == 1 && mode == MG_LEXGS_FORWARD)) continue;
      int* prog = lexgs_progress(nwarps);
      if (!prog) return MG_ERR_UNSUPPORTED;
      if (cudaMemsetAsync(prog, 0, sizeof(int) * nwarps, st) != cudaSuccess) return MG_ERR_LAUNCH;
      if (dir == 0)
        lexgs_pipe_kernel<T, true><<<blocks, 32, 0, st>>>(u, f, nx, ny, ld_u, ld_f, prog, sc);
      else
        lexgs_pipe_kernel<T, false><<<blocks, 32, 0, st>>>(u, f, nx, ny, ld_u, ld_f, prog, sc);
      ++launches;
    }
  return check_launch("mg_smooth_lexgs", launches);
}

extern "C" {

int mg_abi_version(void) { return 1; }

const char* mg_status_string(int status) {
  switch (status) {
    case MG_OK: return "ok";
    case MG_ERR_BADARG: return "bad argument (null pointer, size < 3, ld < ny, ...)";
    case MG_ERR_DTYPE: return "unsupported dtype code";
    case MG_ERR_ALIGN: return "pointer or pitch not aligned for the vector path";
    case MG_ERR_LAUNCH: return mg::g_err[0] ? mg::g_err : "CUDA launch failed";
    case MG_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

int mg_device_sm_count(void) { return mg::sm_count(); }

long long mg_launch_count(void) { return mg::g_launches.load(std::memory_order_relaxed); }

int mg_apply_laplacian(const void* u, void* out, int nx, int ny, int64_t ld_u, int64_t ld_out, double hx,
                       double hy, double coefficient, int dtype, void* stream) {
  MG_REQUIRE(u && out && nx >= 3 && ny >= 3 && ld_u >= ny && ld_out >= ny && hx > 0 && hy > 0);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const dim3 g = grid2d(nx, ny), b(BX, BY);
  if (dtype == MG_F64)
    apply_kernel<double><<<g, b, 0, as_stream(stream)>>>((const double*)u, (double*)out, nx, ny, ld_u, ld_out,
                                                         make_scalars<double>(hx, hy, 1.0, coefficient));
  else
    apply_kernel<float><<<g, b, 0, as_stream(stream)>>>((const float*)u, (float*)out, nx, ny, ld_u, ld_out,
                                                        make_scalars<float>(hx, hy, 1.0, coefficient));
  return check_launch("mg_apply_laplacian");
}

int mg_residual(const void* u, const void* f, void* r, int nx, int ny, int64_t ld_u, int64_t ld_f, int64_t ld_r,
                double hx, double hy, double coefficient, int dtype_in, int dtype_out, void* stream) {
  return mg_residual_h(u, f, r, nx, ny, ld_u, ld_f, ld_r, hx, hy, coefficient, 0.0, dtype_in, dtype_out, stream);
}

int mg_residual_h(const void* u, const void* f, void* r, int nx, int ny, int64_t ld_u, int64_t ld_f, int64_t ld_r,
                  double hx, double hy, double coefficient, double shift, int dtype_in, int dtype_out, void* stream) {
  MG_REQUIRE(u && f && r && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && ld_r >= ny && hx > 0 && hy > 0);
  MG_REQUIRE(shift >= 0.0);
  if (!valid_dtype(dtype_in) || !valid_dtype(dtype_out)) return MG_ERR_DTYPE;
  const dim3 g = grid2d(nx, ny), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  if (dtype_in == MG_F64 && dtype_out == MG_F64)
    residual_kernel<double, double, double><<<g, b, 0, st>>>((const double*)u, (const double*)f, (double*)r, nx, ny,
                                                             ld_u, ld_f, ld_r,
                                                             make_scalars<double>(hx, hy, 1.0, coefficient, shift));
  else if (dtype_in == MG_F32 && dtype_out == MG_F32)
    residual_kernel<float, float, float><<<g, b, 0, st>>>((const float*)u, (const float*)f, (float*)r, nx, ny, ld_u,
                                                          ld_f, ld_r, make_scalars<float>(hx, hy, 1.0, coefficient, shift));
  else if (dtype_in == MG_F32 && dtype_out == MG_F64)
    residual_kernel<float, double, double><<<g, b, 0, st>>>((const float*)u, (const float*)f, (double*)r, nx, ny,
                                                            ld_u, ld_f, ld_r,
                                                            make_scalars<double>(hx, hy, 1.0, coefficient, shift));
  else
    residual_kernel<double, float, double><<<g, b, 0, st>>>((const double*)u, (const double*)f, (float*)r, nx, ny,
                                                            ld_u, ld_f, ld_r,
                                                            make_scalars<double>(hx, hy, 1.0, coefficient, shift));
  return check_launch("mg_residual");
}

int mg_smooth_rbgs(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                   double omega, int sweeps, int dtype, void* stream) {
  return mg_smooth_rbgs_h(u, f, nx, ny, ld_u, ld_f, hx, hy, omega, 0.0, sweeps, dtype, stream);
}

int mg_smooth_rbgs_h(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                     double omega, double shift, int sweeps, int dtype, void* stream) {
  MG_REQUIRE(u && f && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && hx > 0 && hy > 0 && sweeps >= 0);
  MG_REQUIRE(shift >= 0.0);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const int nj = (ny - 2 + 1) / 2;  // max points of one colour in a row
  const dim3 g((nj + BX - 1) / BX, (nx - 2 + BY - 1) / BY), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  for (int k = 0; k < sweeps; ++k)
    for (int colour = 0; colour < 2; ++colour) {
      if (dtype == MG_F64)
        rbgs_colour_kernel<double><<<g, b, 0, st>>>((double*)u, (const double*)f, nx, ny, ld_u, ld_f, colour,
                                                    make_scalars<double>(hx, hy, omega, 1.0, shift));
      else
        rbgs_colour_kernel<float><<<g, b, 0, st>>>((float*)u, (const float*)f, nx, ny, ld_u, ld_f, colour,
                                                   make_scalars<float>(hx, hy, omega, 1.0, shift));
    }
  return check_launch("mg_smooth_rbgs", 2 * sweeps);
}

int mg_smooth_jacobi(void* u, void* tmp, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx,
                     double hy, double omega, int sweeps, int dtype, void* stream) {
  MG_REQUIRE(u && tmp && f && u != tmp && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && hx > 0 && hy > 0 &&
             sweeps >= 0);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const dim3 g = grid2d(nx, ny), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  void* src = u;
  void* dst = tmp;
  for (int k = 0; k < sweeps; ++k) {
    if (dtype == MG_F64)
      jacobi_kernel<double><<<g, b, 0, st>>>((const double*)src, (double*)dst, (const double*)f, nx, ny, ld_u, ld_f,
                                             make_scalars<double>(hx, hy, omega, 1.0));
    else
      jacobi_kernel<float><<<g, b, 0, st>>>((const float*)src, (float*)dst, (const float*)f, nx, ny, ld_u, ld_f,
                                            make_scalars<float>(hx, hy, omega, 1.0));
    void* t = src; src = dst; dst = t;
  }
  if (src != u) {  // odd number of sweeps: result sits in tmp
    const size_t esz = dtype == MG_F64 ? 8 : 4;
    cudaMemcpy2DAsync(u, ld_u * esz, src, ld_u * esz, (size_t)ny * esz, nx, cudaMemcpyDeviceToDevice, st);
  }
  return check_launch("mg_smooth_jacobi", sweeps);
}

int mg_smooth_lexgs(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                    double omega, int sweeps, int mode, int dtype, void* stream) {
  MG_REQUIRE(u && f && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && hx > 0 && hy > 0 && sweeps >= 0);
  MG_REQUIRE(mode >= 0 && mode <= 2);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    return launch_lexgs<double>((double*)u, (const double*)f, nx, ny, ld_u, ld_f, hx, hy, omega, sweeps, mode, st);
  return launch_lexgs<float>((float*)u, (const float*)f, nx, ny, ld_u, ld_f, hx, hy, omega, sweeps, mode, st);
}

int mg_coarse_solve_lexgs(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                          double omega, double coefficient, double tolerance, int max_iterations, double* info,
                          int dtype, void* stream) {
  return mg_coarse_solve_lexgs_h(u, f, nx, ny, ld_u, ld_f, hx, hy, omega, coefficient, 0.0, tolerance, max_iterations,
                                 info, dtype, stream);
}

int mg_coarse_solve_lexgs_h(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                            double omega, double coefficient, double shift, double tolerance, int max_iterations,
                            double* info, int dtype, void* stream) {
  MG_REQUIRE(u && f && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && hx > 0 && hy > 0 && max_iterations >= 1);
  MG_REQUIRE(shift >= 0.0);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const int diag = (nx < ny ? nx : ny) - 2;
  int threads = 32;
  while (threads < diag && threads < 1024) threads <<= 1;
  const size_t esz = dtype == MG_F64 ? 8 : 4;
  const size_t smem = 2 * (size_t)nx * ny * esz;
  const bool use_smem = smem <= 40 * 1024;
  cudaStream_t st = as_stream(stream);
  const double hxhy = hx * hy;
  if (dtype == MG_F64) {
    auto s = make_scalars<double>(hx, hy, omega, coefficient, shift);
    if (use_smem)
      coarse_solve_kernel<double, true><<<1, threads, smem, st>>>((double*)u, (const double*)f, nx, ny, ld_u, ld_f,
                                                                  hxhy, tolerance, max_iterations, info, s);
    else
      coarse_solve_kernel<double, false><<<1, threads, 0, st>>>((double*)u, (const double*)f, nx, ny, ld_u, ld_f,
                                                                hxhy, tolerance, max_iterations, info, s);
  } else {
    auto s = make_scalars<float>(hx, hy, omega, coefficient, shift);
    if (use_smem)
      coarse_solve_kernel<float, true><<<1, threads, smem, st>>>((float*)u, (const float*)f, nx, ny, ld_u, ld_f,
                                                                 hxhy, tolerance, max_iterations, info, s);
    else
      coarse_solve_kernel<float, false><<<1, threads, 0, st>>>((float*)u, (const float*)f, nx, ny, ld_u, ld_f, hxhy,
                                                               tolerance, max_iterations, info, s);
  }
  return check_launch("mg_coarse_solve_lexgs");
}

int mg_restrict(const void* fine, void* coarse, int nxf, int nyf, int64_t ld_f, int64_t ld_c, int method,
                int dtype_in, int dtype_out, void* stream) {
  MG_REQUIRE(fine && coarse && nxf >= 3 && nyf >= 3 && ld_f >= nyf && method >= 0 && method <= 2);
  MG_REQUIRE(((nxf - 1) % 2 == 0) && ((nyf - 1) % 2 == 0));
  if (!valid_dtype(dtype_in) || !valid_dtype(dtype_out)) return MG_ERR_DTYPE;
  const int nxc = (nxf - 1) / 2 + 1, nyc = (nyf - 1) / 2 + 1;
  MG_REQUIRE(ld_c >= nyc);
  const dim3 g = grid2d(nxc, nyc), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  if (dtype_in == MG_F64 && dtype_out == MG_F64)
    restrict_kernel<double, double><<<g, b, 0, st>>>((const double*)fine, (double*)coarse, nxc, nyc, ld_f, ld_c, method);
  else if (dtype_in == MG_F32 && dtype_out == MG_F32)
    restrict_kernel<float, float><<<g, b, 0, st>>>((const float*)fine, (float*)coarse, nxc, nyc, ld_f, ld_c, method);
  else if (dtype_in == MG_F32 && dtype_out == MG_F64)
    restrict_kernel<float, double><<<g, b, 0, st>>>((const float*)fine, (double*)coarse, nxc, nyc, ld_f, ld_c, method);
  else
    restrict_kernel<double, float><<<g, b, 0, st>>>((const double*)fine, (float*)coarse, nxc, nyc, ld_f, ld_c, method);
  return check_launch("mg_restrict");
}

int mg_prolong(const void* coarse, void* fine, int nxc, int nyc, int64_t ld_c, int64_t ld_f, int method, int add,
               int dtype_in, int dtype_out, void* stream) {
  MG_REQUIRE(coarse && fine && nxc >= 2 && nyc >= 2 && ld_c >= nyc && method >= 0 && method <= 1);
  if (!valid_dtype(dtype_in) || !valid_dtype(dtype_out)) return MG_ERR_DTYPE;
  const int nxf = 2 * (nxc - 1) + 1, nyf = 2 * (nyc - 1) + 1;
  MG_REQUIRE(ld_f >= nyf);
  const dim3 g = grid2d(nxf, nyf), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  if (dtype_in == MG_F64 && dtype_out == MG_F64)
    prolong_kernel<double, double><<<g, b, 0, st>>>((const double*)coarse, (double*)fine, nxf, nyf, ld_c, ld_f, method, add);
  else if (dtype_in == MG_F32 && dtype_out == MG_F32)
    prolong_kernel<float, float><<<g, b, 0, st>>>((const float*)coarse, (float*)fine, nxf, nyf, ld_c, ld_f, method, add);
  else if (dtype_in == MG_F32 && dtype_out == MG_F64)
    prolong_kernel<float, double><<<g, b, 0, st>>>((const float*)coarse, (double*)fine, nxf, nyf, ld_c, ld_f, method, add);
  else
    prolong_kernel<double, float><<<g, b, 0, st>>>((const double*)coarse, (float*)fine, nxf, nyf, ld_c, ld_f, method, add);
  return check_launch("mg_prolong");
}

int mg_sumsq_workspace_doubles(void) { return RED_BLOCKS; }

int mg_sumsq(const void* x, int nx, int ny, int64_t ld, int dtype, double* workspace, double* out, void* stream) {
  MG_REQUIRE(x && workspace && out && nx >= 1 && ny >= 1 && ld >= ny);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const int blocks = nx < RED_BLOCKS ? nx : RED_BLOCKS;
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    sumsq_partial_kernel<double><<<blocks, RED_THREADS, 0, st>>>((const double*)x, nx, ny, ld, workspace);
  else
    sumsq_partial_kernel<float><<<blocks, RED_THREADS, 0, st>>>((const float*)x, nx, ny, ld, workspace);
  final_reduce_kernel<false><<<1, RED_THREADS, 0, st>>>(workspace, blocks, out);
  return check_launch("mg_sumsq", 2);
}

int mg_read_doubles(const double* src, double* dst_host, int n, void* stream) {
  MG_REQUIRE(src && dst_host && n >= 1 && n <= 1024);
  cudaStream_t st = as_stream(stream);
  void* alias = nullptr;
  if (cudaHostGetDevicePointer(&alias, dst_host, 0) == cudaSuccess && alias != nullptr) {
    store_doubles_kernel<<<1, 32, 0, st>>>(src, (double*)alias, n);
    return check_launch("mg_read_doubles");
  }
  cudaGetLastError();  // not mapped pinned memory: the copy engine does it
  return cudaMemcpyAsync(dst_host, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st) == cudaSuccess
             ? MG_OK
             : MG_ERR_BADARG;
}

int mg_cast(const void* src, void* dst, int nx, int ny, int64_t ld_src, int64_t ld_dst, int dtype_src,
            int dtype_dst, void* stream) {
  MG_REQUIRE(src && dst && nx >= 1 && ny >= 1 && ld_src >= ny && ld_dst >= ny);
  if (!valid_dtype(dtype_src) || !valid_dtype(dtype_dst)) return MG_ERR_DTYPE;
  const dim3 g = grid2d(nx, ny), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  if (dtype_src == MG_F64 && dtype_dst == MG_F64)
    cast_kernel<double, double><<<g, b, 0, st>>>((const double*)src, (double*)dst, nx, ny, ld_src, ld_dst);
  else if (dtype_src == MG_F32 && dtype_dst == MG_F32)
    cast_kernel<float, float><<<g, b, 0, st>>>((const float*)src, (float*)dst, nx, ny, ld_src, ld_dst);
  else if (dtype_src == MG_F32 && dtype_dst == MG_F64)
    cast_kernel<float, double><<<g, b, 0, st>>>((const float*)src, (double*)dst, nx, ny, ld_src, ld_dst);
  else
    cast_kernel<double, float><<<g, b, 0, st>>>((const double*)src, (float*)dst, nx, ny, ld_src, ld_dst);
  return check_launch("mg_cast");
}

int mg_axpy(double alpha, const void* x, void* y, int nx, int ny, int64_t ld_x, int64_t ld_y, int dtype_x,
            int dtype_y, void* stream) {
  MG_REQUIRE(x && y && nx >= 1 && ny >= 1 && ld_x >= ny && ld_y >= ny);
  if (!valid_dtype(dtype_x) || !valid_dtype(dtype_y)) return MG_ERR_DTYPE;
  const dim3 g = grid2d(nx, ny), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  if (dtype_x == MG_F64 && dtype_y == MG_F64)
    axpy_kernel<double, double><<<g, b, 0, st>>>(alpha, (const double*)x, (double*)y, nx, ny, ld_x, ld_y);
  else if (dtype_x == MG_F32 && dtype_y == MG_F32)
    axpy_kernel<float, float><<<g, b, 0, st>>>((float)alpha, (const float*)x, (float*)y, nx, ny, ld_x, ld_y);
  else if (dtype_x == MG_F32 && dtype_y == MG_F64)
    axpy_kernel<float, double><<<g, b, 0, st>>>(alpha, (const float*)x, (double*)y, nx, ny, ld_x, ld_y);
  else
    axpy_kernel<double, float><<<g, b, 0, st>>>((float)alpha, (const double*)x, (float*)y, nx, ny, ld_x, ld_y);
  return check_launch("mg_axpy");
}

int mg_zero(void* x, int nx, int64_t ld, int dtype, void* stream) {
  MG_REQUIRE(x && nx >= 1 && ld >= 1);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const size_t esz = dtype == MG_F64 ? 8 : 4;
  cudaMemsetAsync(x, 0, (size_t)nx * ld * esz, as_stream(stream));
  return check_launch("mg_zero", 0);
}

int mg_zero_ring(void* x, int nx, int ny, int64_t ld, int first_row, int last_row, int dtype, void* stream) {
  MG_REQUIRE(x && nx >= 1 && ny >= 1 && ld >= ny);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const int n = nx > ny ? nx : ny;
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    zero_ring_kernel<double><<<(n + 255) / 256, 256, 0, st>>>((double*)x, nx, ny, ld, first_row, last_row);
  else
    zero_ring_kernel<float><<<(n + 255) / 256, 256, 0, st>>>((float*)x, nx, ny, ld, first_row, last_row);
  return check_launch("mg_zero_ring");
}

int mg_heat_rhs(const double* u, const double* f1, const double* f0, const double* a, double* rhs, double* sumsq_out,
                double* workspace, int nx, int ny, int64_t ld_u, int64_t ld_f1, int64_t ld_f0, int64_t ld_a, int64_t ld_rhs,
                double hx, double hy, double c_lap, double c_f1, double c_f0, double lam, int zero_first_row,
                int zero_last_row, int norm_row_lo, int norm_row_hi, void* stream) {
  MG_REQUIRE(u && rhs && sumsq_out && workspace && nx >= 3 && ny >= 3 && ld_u >= ny && ld_rhs >= ny && hx > 0 && hy > 0);
  MG_REQUIRE((!f1 || ld_f1 >= ny) && (!f0 || ld_f0 >= ny) && (!a || ld_a >= ny) && rhs != u);
  const int blocks = nx < RED_BLOCKS ? nx : RED_BLOCKS;
  cudaStream_t st = as_stream(stream);
  const double hx2 = pow(hx, 2.0), hy2 = pow(hy, 2.0);
  heat_rhs_kernel<<<blocks, RED_THREADS, 0, st>>>(u, f1, f0, a, rhs, nx, ny, ld_u, ld_f1, ld_f0, ld_a, ld_rhs, 1.0 / hx2,
                                                  1.0 / hy2, c_lap, f1 ? c_f1 : 0.0, f0 ? c_f0 : 0.0, lam, zero_first_row,
                                                  zero_last_row, norm_row_lo < 0 ? 0 : norm_row_lo,
                                                  (norm_row_hi < 0 || norm_row_hi > nx) ? nx : norm_row_hi, workspace);
  final_reduce_kernel<false><<<1, RED_THREADS, 0, st>>>(workspace, blocks, sumsq_out);
  return check_launch("mg_heat_rhs", 2);
}

int mg_fill_sinsin(void* f, int nx, int ny, int64_t ld, double x0, double x1, double y0, double y1,
                   double amplitude, double kx, double ky, int dtype, void* stream) {
  MG_REQUIRE(f && nx >= 2 && ny >= 2 && ld >= ny);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const dim3 g = grid2d(nx, ny), b(BX, BY);
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    fill_sinsin_kernel<double><<<g, b, 0, st>>>((double*)f, nx, ny, ld, x0, x1, y0, y1, amplitude, kx, ky);
  else
    fill_sinsin_kernel<float><<<g, b, 0, st>>>((float*)f, nx, ny, ld, x0, x1, y0, y1, amplitude, kx, ky);
  return check_launch("mg_fill_sinsin");
}

int mg_maxerr_sinsin(const void* u, int nx, int ny, int64_t ld, double x0, double x1, double y0, double y1,
                     double amplitude, double kx, double ky, int dtype, double* workspace, double* out,
                     void* stream) {
  MG_REQUIRE(u && workspace && out && nx >= 2 && ny >= 2 && ld >= ny);
  if (!valid_dtype(dtype)) return MG_ERR_DTYPE;
  const int blocks = nx < RED_BLOCKS ? nx : RED_BLOCKS;
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    maxerr_partial_kernel<double><<<blocks, RED_THREADS, 0, st>>>((const double*)u, nx, ny, ld, x0, x1, y0, y1,
                                                                  amplitude, kx, ky, workspace);
  else
    maxerr_partial_kernel<float><<<blocks, RED_THREADS, 0, st>>>((const float*)u, nx, ny, ld, x0, x1, y0, y1,
                                                                 amplitude, kx, ky, workspace);
  final_reduce_kernel<true><<<1, RED_THREADS, 0, st>>>(workspace, blocks, out);
  return check_launch("mg_maxerr_sinsin", 2);
}

}  // extern "C"
