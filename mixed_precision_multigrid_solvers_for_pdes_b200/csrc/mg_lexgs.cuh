// Pipelined lexicographic Gauss-Seidel (skewed wavefront over warps): kernel + progress-counter pool, shared by
// mg_smooth_lexgs (mg_basic.cu) and the CorrectedMultigridSolver kernels (mg_corrected.cu).
#pragma once
#include <mutex>
#include "mg_common.cuh"

namespace mg {

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lexicographic GS on any grid size, pipelined over warps ("skewed wavefront").  Warp w owns the 32 interior rows
// 32w .. 32w+31 (in sweep order), lane l one of them; at step t lane l relaxes column t - l, so within a warp the row
// above is always one step ahead (its new value arrives by shuffle) and the row below one step behind (its old value
// too).  Across warps, the last row of warp w-1 must be ahead of the first row of warp w: warp w-1 publishes the
// number of columns its last row has completed in progress[w-1] (st.release.gpu, every 32 steps) and warp w polls it
// (ld.acquire.gpu) before it loads the next 32 values of that row (ld.global.cg).  Dependencies only point to LOWER warp
// indices, i.e. to blocks dispatched earlier, so waiting cannot deadlock whatever the grid size; a 2 s watchdog traps
// instead of hanging should that assumption ever break.  Every point is relaxed by relax_strict with the operands of
// the sequential loop (new above / left, old below / right), so the result is the reference's
// (smoothers.py:153-173) bit for bit; FWD=false runs the mirrored sweep (bottom-right to top-left).
// Cost: ~(ny + 64 * rows/32) steps; a step is ~30 dependent instructions of ONE warp (~0.17 us, ncu: 11 cycles per
// issued instruction), so a sweep takes 0.6 ms at 1025^2, 2.8 ms at 4097^2 and 12 ms at 16385^2
// (profiles/r02_lexgs_bench.log) where the one-block wavefront needed a block barrier per anti-diagonal and 1024
// threads for diagonals of up to 16383 points (~0.6 ms / ~10 ms / > 150 ms).  The ordering itself is the limit: a
// red-black sweep of 16385^2 runs in 0.25 ms at the HBM roofline.
template <typename T, bool FWD, bool CORRECTED = false>
__global__ void __launch_bounds__(32) lexgs_pipe_kernel(T* __restrict__ u, const T* __restrict__ f, int nx, int ny,
                                                         int64_t ldu, int64_t ldf, int* progress, StencilScalars<T> s) {
  // one warp per CTA: the per-row (uncoalesced) loads and stores of a warp keep one SM's load/store unit busy for
  // about as long as the dependent chain of a step lasts, so warps are spread over as many SMs as there are
  const int lane = threadIdx.x, w = blockIdx.x;
  const int nrows = nx - 2, ncols = ny - 2;
  const int r = w * 32 + lane;
  const bool row_ok = r < nrows;
  const int last_lr = min(w * 32 + 31, nrows - 1), last_lane = last_lr - w * 32;
  // logical (sweep-order) row / column -> storage index; -1 and nrows / ncols land on the boundary ring
  auto I = [&](int lr) { return FWD ? 1 + lr : nx - 2 - lr; };
  auto J = [&](int lc) { return FWD ? 1 + lc : ny - 2 - lc; };
  T* urow = u + (int64_t)I(row_ok ? r : last_lr) * ldu;
  const T* frow = f + (int64_t)I(row_ok ? r : last_lr) * ldf;
  const T* uprev = u + (int64_t)I(w * 32 - 1) * ldu;   // row relaxed just before our first one (ring row if w == 0)
  const T* unext = u + (int64_t)I(last_lr + 1) * ldu;  // row relaxed just after our last one (still old)
  const int nsteps = ncols + last_lane;
  T prev = urow[J(-1)];  // value "behind" the current column: the ring value first, then the lane's last result
  T carry = urow[J(0)];  // old value of the current column
  int cached = 0;
  // old values one column ahead (un) and right-hand sides (fc) of 8 steps, double buffered: the loads of the next 8
  // steps are in flight while the current 8 are relaxed
  T un[2][8], fc[2][8];
  auto load8 = [&](int buf, int tc) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = tc + k - lane;
      un[buf][k] = (c + 1 >= 0 && c + 1 <= ncols) ? urow[J(c + 1)] : (T)0;
      fc[buf][k] = (c >= 0 && c < ncols) ? frow[J(c)] : (T)0;
    }
  };
  load8(0, 0);
  for (int tb = 0; tb < nsteps; tb += 32) {
    // once per 32 steps: wait for the row above our first one, fetch 32 of its new values and 32 old values of the
    // row below our last one (one per lane, handed to lane 0 / the last lane by shuffle)
    if (w > 0) {
      const int need = min(tb + 32, ncols);
      if (cached < need) {
        if (lane == 0) {
          const long long t0 = clock64();
          // spin on a relaxed load (an acquire load invalidates this SM's L1 on every iteration), acquire once
          while (ld_relaxed_gpu(progress + w - 1) < need)
            if (clock64() - t0 > 4000000000LL) __trap();
          cached = ld_acquire_gpu(progress + w - 1);
        }
        cached = __shfl_sync(0xffffffffu, cached, 0);
      }
    }
    T up32 = (T)0, dn32 = (T)0;
    if (tb + lane < ncols) up32 = __ldcg(uprev + J(tb + lane));
    {
      const int c = tb + lane - last_lane;
      if (c >= 0 && c < ncols) dn32 = __ldcg(unext + J(c));
    }
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int tc = tb + 8 * sub;
      if (tc >= nsteps) break;  // warp-uniform
      constexpr int NB[4] = {1, 0, 1, 0};
      const int cur = sub & 1;
      load8(NB[sub], tc + 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = tc + k - lane;
        const bool active = row_ok && c >= 0 && c < ncols;
        T before = __shfl_up_sync(0xffffffffu, prev, 1);   // row above (sweep order): new value of column c
        const T b0 = __shfl_sync(0xffffffffu, up32, 8 * sub + k);
        if (lane == 0) before = b0;
        T after = __shfl_down_sync(0xffffffffu, un[cur][k], 1);  // row below: its old value of column c
        const T a0 = __shfl_sync(0xffffffffu, dn32, 8 * sub + k);
        if (lane == last_lane) after = a0;
        T res;
        if constexpr (CORRECTED) {
          // corrected_multigrid.py:263-270: 0.25 * (u[i-1,j] + u[i+1,j] + u[i,j-1] + u[i,j+1] + h^2 f), summed left to right
          using A = Strict<T>;
          res = A::mul((T)0.25, A::add(A::add(A::add(A::add(before, after), prev), un[cur][k]), A::mul(s.hx2, fc[cur][k])));
        } else {
          res = relax_strict<T>(s, carry, after, before, un[cur][k], prev, fc[cur][k]);
        }
        if (active) {
          urow[J(c)] = res;
          prev = res;
          carry = un[cur][k];
        }
      }
    }
    if (lane == last_lane) {
      const int done = min(tb + 32 - last_lane, ncols);
      if (done > 0) st_release_gpu(progress + w, done);
    }
  }
}

// per-device progress counters of lexgs_pipe_kernel: 8 slots used round robin (calls on different streams, and the
// nodes of different captured graphs, get different slots), one int per 32 rows
constexpr int LEXGS_SLOTS = 8, LEXGS_SLOT_INTS = 1 << 15;
inline int* lexgs_progress(int nwarps) {
  static int* buf[64] = {nullptr};
  static unsigned next[64] = {0};
  static std::mutex guard;  // host threads may call the entry points concurrently
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || nwarps > LEXGS_SLOT_INTS) return nullptr;
  std::lock_guard<std::mutex> lock(guard);
  if (!buf[dev] && cudaMalloc(&buf[dev], sizeof(int) * LEXGS_SLOTS * LEXGS_SLOT_INTS) != cudaSuccess) return nullptr;
  return buf[dev] + (size_t)(next[dev]++ % LEXGS_SLOTS) * LEXGS_SLOT_INTS;
}

}  // namespace mg
