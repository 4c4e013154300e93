// instantiation unit of the CTA-cooperative stream kernel for 2 spline dimension(s)
#include "stream_cta.cuh"
namespace gwi {
stream_fn pick_stream_cta_ns2(int nd, int nlin) { return pick_stream_cta_for_ns<2>(nd, nlin); }
}  // namespace gwi
