// stream kernel instantiations for 3 spline dimension(s) (one translation unit per count so that
// make -j compiles them in parallel)
#include "stream.cuh"
namespace gwi {
stream_fn pick_stream_ns3(int nd, int nlin, bool g2, bool param, bool maxonly) { return pick_stream_for_ns<3>(nd, nlin, g2, param, maxonly); }
}  // namespace gwi
