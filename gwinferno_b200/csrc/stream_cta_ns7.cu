// instantiation unit of the CTA-cooperative stream kernel for 7 spline dimension(s)
#include "stream_cta.cuh"
namespace gwi {
stream_fn pick_stream_cta_ns7(int nd, int nlin) { return pick_stream_cta_for_ns<7>(nd, nlin); }
}  // namespace gwi
