// instantiation unit of the CTA-cooperative stream kernel for 6 spline dimension(s)
#include "stream_cta.cuh"
namespace gwi {
stream_fn pick_stream_cta_ns6(int nd, int nlin) { return pick_stream_cta_for_ns<6>(nd, nlin); }
}  // namespace gwi
