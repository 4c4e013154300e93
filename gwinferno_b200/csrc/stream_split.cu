// instantiations of the role-split experiment kernel (empty unless built with -DGWI_EXP_SPLIT=1)
#include "stream_split.cuh"

namespace gwi {
#if GWI_EXP_SPLIT
template <int NS>
static stream_fn pick_split_ns(int nd, int nlin, bool g2) {
#define GWI_SPLIT_PICK(ND, NL)                                                                                   \
  if (nd == ND && nlin == NL) return g2 ? (stream_fn)stream_split_kernel<NS, (ND <= NS ? ND : 0), NL, true> : (stream_fn)stream_split_kernel<NS, (ND <= NS ? ND : 0), NL, false>;
  GWI_SPLIT_PICK(0, 1)
  GWI_SPLIT_PICK(1, 1)
  GWI_SPLIT_PICK(2, 1)
  GWI_SPLIT_PICK(3, 1)
  GWI_SPLIT_PICK(4, 1)
  GWI_SPLIT_PICK(3, 2)
  GWI_SPLIT_PICK(4, 2)
#undef GWI_SPLIT_PICK
  return nullptr;
}
stream_fn pick_stream_split_kernel(int ns, int ndeep, int nlin, bool g2) {
  if (ns == 7) return pick_split_ns<7>(ndeep, nlin, g2);  // the cfg-2/3 model family
  if (ns == 5) return pick_split_ns<5>(ndeep, nlin, g2);  // IID spins (cfg 5)
  return nullptr;
}
size_t stream_split_extra_smem(int npairs) { return (size_t)npairs * (SPLIT_RING * 64 * sizeof(double) + sizeof(SplitSync)); }
#else
typedef void (*stream_fn)(const ModelDev*);
stream_fn pick_stream_split_kernel(int, int, int, bool) { return nullptr; }
size_t stream_split_extra_smem(int) { return 0; }
#endif
}  // namespace gwi
