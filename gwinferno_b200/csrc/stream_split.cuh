// EXPERIMENT (compiled only with -DGWI_EXP_SPLIT=1, selected at run time with GWI_SPLIT=1; never part
// of the default build): role-split variant of the stream kernel.
//
// The stream kernel is bound by issue latency at 8 warps/SM because every warp keeps the coefficient
// registers of the leading dims, their moment registers, two samples in flight and a load buffer:
// 255 registers per thread.  Here a block holds warp PAIRS:
//   producer warp  -- coefficients of the leading dims in registers, loads the plan words, evaluates the
//                     cubics and the exp, and hands p = exp(x - shift) to its partner through a small
//                     shared-memory ring (512 B per iteration of 64 samples);
//   consumer warp  -- re-reads the words of the same iteration (L2 hits), takes p from the ring and does all
//                     accumulation: S1, S2, linear-term gradients, register moments of the leading dims
//                     with their spills, the lane-pair-private deep accumulators, the record flush.
// Each role needs about half of the state, so 16 warps fit where 8 did (launch bound 512 threads =>
// <= 128 registers), for ~20 % more issued instructions (the words are loaded and unpacked twice).
// Producer and consumer walk the same slices (the producer pulls them from the global counter and passes
// them on), so records and reduction tree are those of stream_kernel and the sums are the same up to the
// order of the rare shared-memory spill atomics.
// Restrictions: spline + register-resident linear terms only (no generic-term loop), full pass only.
#pragma once
#include "stream.cuh"

#ifndef GWI_EXP_SPLIT
#define GWI_EXP_SPLIT 0
#endif

#if GWI_EXP_SPLIT
namespace gwi {

constexpr int SPLIT_RING = 4;  // iterations in flight between producer and consumer

// spin-wait hint: the host emulator must hand the processor to the partner fiber
__device__ __forceinline__ void split_spin() {
#ifdef GWI_HOST_EMULATION
  gwi_emu::warp_yield();
#else
  __nanosleep(20);
#endif
}
__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

struct SplitSync {  // one per pair, in shared memory (32 bytes)
  int produced;     // iterations whose p values are in the ring
  int consumed;     // iterations the consumer has taken out
  int slice_seq;    // slices announced by the producer
  int slice_id;     // the announced slice (>= n_slices: stop)
  int slice_ack;    // slices the consumer has read (the mailbox holds one announcement)
  int pad[3];
};

template <int NS, int NDEEP, int NLIN, bool G2>
__global__ void __launch_bounds__(512, 1) stream_split_kernel(const ModelDev* __restrict__ Mp) {
  const ModelDev& M = Mp[blockIdx.y];
  constexpr int NSH = NS - NDEEP;
  constexpr int MOM = G2 ? 2 : 1;
  constexpr int NSd = NS > 0 ? NS : 1;
  constexpr int NSHd = NSH > 0 ? NSH : 1;
  constexpr int NLd = NLIN > 0 ? NLIN : 1;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, npairs = blockDim.x >> 6;
  const bool producer = warp < npairs;
  const int pair = producer ? warp : warp - npairs;
  const int rows_total = M.rows_total;
  const int n_kops = M.n_kops, n_gs = M.n_gslots;
  // shared layout (doubles): tables | kc | kops copy | per pair { msh | deep | (unused generic scratch) } |
  //                          ring[npairs][SPLIT_RING][64] | sync[npairs]
  double* tables = sm;
  double* kcs = tables + rows_total * 4;
  KopDev* kops_s = reinterpret_cast<KopDev*>(kcs + n_kops * KC_STRIDE);
  double* wbase = reinterpret_cast<double*>(kops_s + n_kops);
  const int per_warp = rows_total * 4 * MOM + M.deep_entries * 2 * DEEP_LANES + n_gs * 32 * (1 + MOM);
  double* msh = wbase + (size_t)pair * per_warp;
  double2* deep = reinterpret_cast<double2*>(msh + rows_total * 4 * MOM);
  double* ring = wbase + (size_t)npairs * per_warp + (size_t)pair * SPLIT_RING * 64;
  SplitSync* sync = reinterpret_cast<SplitSync*>(wbase + (size_t)npairs * per_warp + (size_t)npairs * SPLIT_RING * 64) + pair;
  for (int i = threadIdx.x; i < rows_total * 4; i += blockDim.x) tables[i] = M.tables[i];
  for (int i = threadIdx.x; i < n_kops * KC_STRIDE; i += blockDim.x) kcs[i] = M.kc[i];
  for (int i = threadIdx.x; i < n_kops; i += blockDim.x) kops_s[i] = M.kops[i];
  if (!producer)
    for (int i = lane; i < per_warp; i += 32) msh[i] = 0.0;
  if (producer && lane == 0) {
    sync->produced = 0;
    sync->consumed = 0;
    sync->slice_seq = 0;
    sync->slice_id = 0;
    sync->slice_ack = 0;
  }
  __syncthreads();

  const double* tab_d[NSd];
  int row_off[NSd], rows_d[NSd], deep_off[NSd];
#pragma unroll
  for (int d = 0; d < NS; ++d) {
    row_off[d] = M.dims[d].row_off;
    rows_d[d] = M.dims[d].rows;
    deep_off[d] = M.dims[d].deep_off;
    tab_d[d] = tables + row_off[d] * 4;
  }
  int lin_col[NLd];
#pragma unroll
  for (int l = 0; l < NLIN; ++l) lin_col[l] = kops_s[l].col0;
  const int ncol = M.n_columns;
  const size_t blk_words = (size_t)ncol * 64;
  const uint64_t* __restrict__ cols = M.columns;
  const int col_static = M.col_static;
  int k_iter = 0;     // iterations handed over so far (ring position), identical in both warps of the pair
  int slices_seen = 0;

  if (producer) {
    // =========================== producer: p for every sample ===========================
    double theta[NLd];
#pragma unroll
    for (int l = 0; l < NLIN; ++l) theta[l] = kcs[l * KC_STRIDE];
    int cur[NSHd];
    double cf[NSHd][4];
#pragma unroll
    for (int d = 0; d < NSH; ++d) {
      cur[d] = -1;
#pragma unroll
      for (int n = 0; n < 4; ++n) cf[d][n] = 0.0;
    }
    for (;;) {
      int sl = 0;
      if (lane == 0) {
        sl = atomicAdd(M.slice_counter, 1);
        while (ld_volatile(&sync->slice_ack) < slices_seen) split_spin();  // the previous announcement has been read
        sync->slice_id = sl;
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&sync->slice_seq) = ++slices_seen;
      }
      sl = __shfl_sync(0xffffffffu, sl, 0);
      if (sl >= M.n_slices) break;
      for (int c = M.slice_begin[sl]; c < M.slice_begin[sl + 1]; ++c) {
        const Chunk C = M.chunks[c];
        const double shift = M.shift[C.segment];
        const uint64_t* __restrict__ cbase = cols + (size_t)(C.first >> 6) * blk_words + lane * UNROLL;
        const int iters = C.steps / UNROLL;
        for (int it = 0; it < iters; ++it) {
          const uint64_t* q = cbase + (size_t)it * blk_words;
          ulonglong2 w[NSd];
#pragma unroll
          for (int d = 0; d < NS; ++d) w[d] = __ldg(reinterpret_cast<const ulonglong2*>(q + d * 64));
          const double2 st = __ldg(reinterpret_cast<const double2*>(q + col_static * 64));
          double2 lin[NLd];
#pragma unroll
          for (int l = 0; l < NLIN; ++l) lin[l] = __ldg(reinterpret_cast<const double2*>(q + lin_col[l] * 64));
          double p2[2];
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            double x = s == 0 ? st.x : st.y;
#pragma unroll
            for (int d = 0; d < NS; ++d) {
              const unsigned long long word = s == 0 ? w[d].x : w[d].y;
              const int hi = (int)(word >> 32);
              const int J = (unsigned)hi >> 20;
              const double wv = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, (int)(unsigned)word) - 1.5;
              double v;
              if (d < NSH) {
                if (J != cur[d]) {  // the producer only follows the coefficients; moments live in the consumer
                  cur[d] = J;
                  const double2 a01 = *reinterpret_cast<const double2*>(tab_d[d] + J * 4);
                  const double2 a23 = *reinterpret_cast<const double2*>(tab_d[d] + J * 4 + 2);
                  cf[d][0] = a01.x;
                  cf[d][1] = a01.y;
                  cf[d][2] = a23.x;
                  cf[d][3] = a23.y;
                }
                v = fma(fma(cf[d][3], wv, cf[d][2]), wv * wv, fma(cf[d][1], wv, cf[d][0]));
              } else {
                const double2 a01 = *reinterpret_cast<const double2*>(tab_d[d] + J * 4);
                const double2 a23 = *reinterpret_cast<const double2*>(tab_d[d] + J * 4 + 2);
                v = fma(fma(a23.y, wv, a23.x), wv * wv, fma(a01.y, wv, a01.x));
              }
              x += v;
            }
#pragma unroll
            for (int l = 0; l < NLIN; ++l) x = fma(theta[l], s == 0 ? lin[l].x : lin[l].y, x);
            p2[s] = exp_nonpos(x - shift);
          }
          // ring slot free?  (the consumer has taken iteration k_iter - SPLIT_RING out)
          while (k_iter - ld_volatile(&sync->consumed) >= SPLIT_RING) split_spin();
          *reinterpret_cast<double2*>(ring + (k_iter % SPLIT_RING) * 64 + lane * 2) = make_double2(p2[0], p2[1]);
          __syncwarp();
          ++k_iter;
          if (lane == 0) {
            __threadfence_block();
            *reinterpret_cast<volatile int*>(&sync->produced) = k_iter;
          }
        }
      }
    }
    return;
  }

  // =========================== consumer: all accumulation ===========================
  double2* deep_d[NSd];
#pragma unroll
  for (int d = 0; d < NS; ++d) deep_d[d] = deep + (size_t)deep_off[d] * DEEP_LANES + (lane & (DEEP_LANES - 1));
  const int m2_off = rows_total * 4;
  double S1 = 0.0, S2 = 0.0;
  int cur[NSHd];
  double m1[NSHd][4];
  double m2[(G2 && NSH > 0) ? NSH : 1][4];
#pragma unroll
  for (int d = 0; d < NSH; ++d) {
    cur[d] = -1;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      m1[d][n] = 0.0;
      if (G2) m2[d][n] = 0.0;
    }
  }
  double gl1[NLd], gl2[NLd];
#pragma unroll
  for (int l = 0; l < NLIN; ++l) {
    gl1[l] = 0.0;
    gl2[l] = 0.0;
  }
  for (;;) {
    ++slices_seen;
    while (ld_volatile(&sync->slice_seq) < slices_seen) split_spin();
    __threadfence_block();
    const int sl = ld_volatile(&sync->slice_id);
    __syncwarp();  // every lane has read the announcement
    if (lane == 0) *reinterpret_cast<volatile int*>(&sync->slice_ack) = slices_seen;
    if (sl >= M.n_slices) break;
    for (int c = M.slice_begin[sl]; c < M.slice_begin[sl + 1]; ++c) {
      const Chunk C = M.chunks[c];
      const uint64_t* __restrict__ cbase = cols + (size_t)(C.first >> 6) * blk_words + lane * UNROLL;
      const int iters = C.steps / UNROLL;
      for (int it = 0; it < iters; ++it) {
        const uint64_t* q = cbase + (size_t)it * blk_words;
        ulonglong2 w[NSd];
#pragma unroll
        for (int d = 0; d < NS; ++d) w[d] = __ldg(reinterpret_cast<const ulonglong2*>(q + d * 64));
        double2 lin[NLd];
#pragma unroll
        for (int l = 0; l < NLIN; ++l) lin[l] = __ldg(reinterpret_cast<const double2*>(q + lin_col[l] * 64));
        while (ld_volatile(&sync->produced) <= k_iter) split_spin();
        __threadfence_block();
        const double2 pp = *reinterpret_cast<const double2*>(ring + (k_iter % SPLIT_RING) * 64 + lane * 2);
        __syncwarp();
        ++k_iter;
        if (lane == 0) *reinterpret_cast<volatile int*>(&sync->consumed) = k_iter;
        int Jd[2][NSd];
        double wd[2][NSd];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double p = s == 0 ? pp.x : pp.y, p2 = p * p;
          S1 += p;
          S2 += p2;
#pragma unroll
          for (int l = 0; l < NLIN; ++l) {
            const double f = s == 0 ? lin[l].x : lin[l].y;
            gl1[l] = fma(p, f, gl1[l]);
            if (G2) gl2[l] = fma(p2, f, gl2[l]);
          }
#pragma unroll
          for (int d = 0; d < NS; ++d) {
            const unsigned long long word = s == 0 ? w[d].x : w[d].y;
            const int hi = (int)(word >> 32);
            const int J = (unsigned)hi >> 20;
            const double wv = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, (int)(unsigned)word) - 1.5;
            Jd[s][d] = J;
            wd[s][d] = wv;
            if (d < NSH) {
              if (J != cur[d]) {
                if (cur[d] >= 0) spill_moments<G2>(msh, (row_off[d] + cur[d]) * 4, m2_off, m1[d], m2[G2 ? d : 0]);
                cur[d] = J;
              }
              const double w2 = wv * wv, w3 = w2 * wv;
              m1[d][0] += p;
              m1[d][1] = fma(p, wv, m1[d][1]);
              m1[d][2] = fma(p, w2, m1[d][2]);
              m1[d][3] = fma(p, w3, m1[d][3]);
              if (G2) {
                m2[d][0] += p2;
                m2[d][1] = fma(p2, wv, m2[d][1]);
                m2[d][2] = fma(p2, w2, m2[d][2]);
                m2[d][3] = fma(p2, w3, m2[d][3]);
              }
            }
          }
        }
        if (NDEEP > 0) {
          // lanes 0-15 update the pair-shared accumulators first, then lanes 16-31 (as in stream_kernel)
#pragma unroll
          for (int half = 0; half < (DEEP_LANES == 32 ? 1 : 2); ++half) {
            if (DEEP_LANES == 32 || (lane >> 4) == half) {
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                const double p = s == 0 ? pp.x : pp.y;
#pragma unroll
                for (int d = NSH; d < NS; ++d) {
                  const double wv = wd[s][d], w2 = wv * wv, w3 = w2 * wv;
                  double2* e = deep_d[d] + (size_t)Jd[s][d] * (2 * MOM * DEEP_LANES);
                  double2 v0 = e[0], v1 = e[DEEP_LANES];
                  v0.x += p;
                  v0.y = fma(p, wv, v0.y);
                  v1.x = fma(p, w2, v1.x);
                  v1.y = fma(p, w3, v1.y);
                  e[0] = v0;
                  e[DEEP_LANES] = v1;
                  if (G2) {
                    const double p2 = p * p;
                    double2 u0 = e[2 * DEEP_LANES], u1 = e[3 * DEEP_LANES];
                    u0.x += p2;
                    u0.y = fma(p2, wv, u0.y);
                    u1.x = fma(p2, w2, u1.x);
                    u1.y = fma(p2, w3, u1.y);
                    e[2 * DEEP_LANES] = u0;
                    e[3 * DEEP_LANES] = u1;
                  }
                }
              }
            }
            __syncwarp();
          }
        }
      }
      // ---- write this chunk's record and clear the accumulators (same as stream_kernel) ----
#pragma unroll
      for (int d = 0; d < NSH; ++d) flush_moments<G2>(msh, cur[d] >= 0 ? row_off[d] + cur[d] : -1, m2_off, lane, m1[d], m2[G2 ? d : 0]);
      __syncwarp();
      double* rec = M.records0 + (size_t)C.record_slot * M.rec_doubles;
      const double s1 = wsum(S1), s2 = wsum(S2);
      S1 = 0.0;
      S2 = 0.0;
      if (lane == 0) {
        rec[0] = s1;
        rec[1] = s2;
      }
#pragma unroll
      for (int l = 0; l < NLIN; ++l) {
        const double a = wsum(gl1[l]), b = wsum(gl2[l]);
        gl1[l] = 0.0;
        gl2[l] = 0.0;
        if (lane == 0) {
          rec[2 + l] = a;
          if (G2) rec[2 + n_gs + l] = b;
        }
      }
#pragma unroll
      for (int d = NSH; d < NS; ++d) {
        const int ne = rows_d[d] * 2 * MOM;
        for (int e = lane; e < ne; e += 32) {
          double2* row = deep + (size_t)(deep_off[d] + e) * DEEP_LANES;
          double ax = 0.0, ay = 0.0;
          for (int i = 0; i < DEEP_LANES; ++i) {
            const int l = (i + lane) & (DEEP_LANES - 1);
            const double2 v = row[l];
            ax += v.x;
            ay += v.y;
            row[l] = make_double2(0.0, 0.0);
          }
          const int J = e / (2 * MOM), r = e - J * 2 * MOM, mm = r >> 1, pr = r & 1;
          const int o = mm * m2_off + (row_off[d] + J) * 4 + pr * 2;
          msh[o] = ax;
          msh[o + 1] = ay;
        }
      }
      __syncwarp();
      double* recM = rec + 2 + n_gs * MOM;
      for (int i = lane; i < rows_total * 4 * MOM; i += 32) {
        recM[i] = msh[i];
        msh[i] = 0.0;
      }
      __syncwarp();
    }
  }
}

// nullptr when this (dims, deep dims, linear terms) combination is not instantiated: the caller keeps stream_kernel
stream_fn pick_stream_split_kernel(int ns, int ndeep, int nlin, bool g2);
size_t stream_split_extra_smem(int npairs);

}  // namespace gwi
#endif  // GWI_EXP_SPLIT
