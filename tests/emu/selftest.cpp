// TEST INFRASTRUCTURE -- self-test kernels of the host warp emulator (tests/test_emulator_runtime.py):
// every intrinsic the emulator offers, with results that are known in closed form.
#include <cuda_runtime.h>

#include <cstdio>

// `extern __shared__ double sm[]` names gwi::sm of emu_runtime.cpp, as in the product kernels
namespace gwi {

__global__ void k_shuffles(double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v = (double)(threadIdx.x + 1);
  double s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);  // warp sum
  const double up = __shfl_up_sync(0xffffffffu, v, 3), down = __shfl_down_sync(0xffffffffu, v, 5), bc = __shfl_sync(0xffffffffu, v, 7);
  const unsigned bal = __ballot_sync(0xffffffffu, lane % 3 == 0);
  const int any = __any_sync(0xffffffffu, lane == 31), all = __all_sync(0xffffffffu, lane < 32);
  const unsigned same = __match_any_sync(0xffffffffu, lane / 8);
  double* o = out + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 8;
  o[0] = s;
  o[1] = up;
  o[2] = down;
  o[3] = bc;
  o[4] = (double)bal;
  o[5] = (double)(any + 2 * all);
  o[6] = (double)same;
  o[7] = (double)warp;
}

// block reduction through static + dynamic shared memory, early exit of half of the threads
__global__ void k_block(const double* in, double* out, int n) {
  extern __shared__ double sm[];
  __shared__ double total;
  if (threadIdx.x == 0) total = 0.0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  sm[threadIdx.x] = i < n ? in[i] : 0.0;
  __syncthreads();
  if (threadIdx.x >= blockDim.x / 2) return;  // exited threads no longer take part in barriers
  sm[threadIdx.x] += sm[threadIdx.x + blockDim.x / 2];
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = 0.0;
    for (unsigned k = threadIdx.x; k < blockDim.x / 2; k += 32) v += sm[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) total = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(out, total);  // across blocks (OS threads): a real atomic
}

// two-phase update of a slot shared by lanes l and l+16 (the stream kernel's deep accumulators),
// partial-mask __syncwarp inside a divergent branch, named barriers between warp pairs
__global__ void k_sync(double* out, int* counters) {
  __shared__ double slot[8][16];
  __shared__ double mailbox[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < 16) slot[warp][lane] = 0.0;
  __syncwarp();
  for (int half = 0; half < 2; ++half) {
    if ((lane >> 4) == half) slot[warp][lane & 15] += (double)(lane + 1);
    __syncwarp();
  }
  if (lane < 8) {  // only 8 lanes take part
    slot[warp][lane] += 100.0;
    __syncwarp(0x000000ffu);
    if (lane == 0) {
      double s = 0.0;
      for (int i = 0; i < 8; ++i) s += slot[warp][i];
      out[blockIdx.x * 8 + warp] = s;
    }
  }
  // producer (even warp) / consumer (odd warp) pairs hand a value over with bar.sync id, 64
  const int pair = warp >> 1;
  if ((warp & 1) == 0) {
    if (lane == 0) mailbox[pair] = 1000.0 + warp;
    gwi_named_barrier_sync(1 + pair, 64);
  } else {
    gwi_named_barrier_sync(1 + pair, 64);
    if (lane == 0) out[64 + blockIdx.x * 4 + pair] = mailbox[pair];
  }
  if (lane == 0) {
    atomicMax(&counters[0], (int)(blockIdx.x * 8 + warp));
    atomicMin(&counters[1], -(int)(blockIdx.x * 8 + warp));
    atomicAdd(&counters[2], 1);
  }
}

}  // namespace gwi
using namespace gwi;

extern "C" int gwi_emu_selftest(double* shuffles /*[2*64*8]*/, const double* in, int n, double* block_sum, double* sync_out /*[64 + 16]*/, int* counters /*[3]*/) {
  GWI_EMU_LAUNCH(k_shuffles, 2, 64, 0, 0)(shuffles);
  *block_sum = 0.0;
  GWI_EMU_LAUNCH(k_block, (n + 255) / 256, 256, 256 * sizeof(double), 0)(in, block_sum, n);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  // the same launch recorded in a stream capture and replayed twice
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  GWI_EMU_LAUNCH(k_sync, 4, 256, 0, st)(sync_out, counters);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  cudaGraphLaunch(ge, st);
  cudaGraphLaunch(ge, st);
  cudaGraphExecDestroy(ge);
  cudaStreamDestroy(st);
  return 0;
}
