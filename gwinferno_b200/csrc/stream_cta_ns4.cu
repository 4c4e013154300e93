// instantiation unit of the CTA-cooperative stream kernel for 4 spline dimension(s)
#include "stream_cta.cuh"
namespace gwi {
stream_fn pick_stream_cta_ns4(int nd, int nlin) { return pick_stream_cta_for_ns<4>(nd, nlin); }
}  // namespace gwi
