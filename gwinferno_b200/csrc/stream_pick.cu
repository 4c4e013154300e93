// dispatch to the per-dimension-count instantiation units
#include "stream_cta.cuh"
namespace gwi {
stream_fn pick_stream_ns0(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns1(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns2(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns3(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns4(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns5(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns6(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns7(int nd, int nlin, bool g2, bool param, bool maxonly);
stream_fn pick_stream_ns8(int nd, int nlin, bool g2, bool param, bool maxonly);

stream_fn pick_stream_kernel(int ns, int ndeep, int nlin, bool g2, bool param, bool maxonly) {
  switch (ns) {
    case 0: return pick_stream_ns0(ndeep, nlin, g2, param, maxonly);
    case 1: return pick_stream_ns1(ndeep, nlin, g2, param, maxonly);
    case 2: return pick_stream_ns2(ndeep, nlin, g2, param, maxonly);
    case 3: return pick_stream_ns3(ndeep, nlin, g2, param, maxonly);
    case 4: return pick_stream_ns4(ndeep, nlin, g2, param, maxonly);
    case 5: return pick_stream_ns5(ndeep, nlin, g2, param, maxonly);
    case 6: return pick_stream_ns6(ndeep, nlin, g2, param, maxonly);
    case 7: return pick_stream_ns7(ndeep, nlin, g2, param, maxonly);
    case 8: return pick_stream_ns8(ndeep, nlin, g2, param, maxonly);
  }
  return nullptr;
}

stream_fn pick_stream_cta_ns1(int nd, int nlin);
stream_fn pick_stream_cta_ns2(int nd, int nlin);
stream_fn pick_stream_cta_ns3(int nd, int nlin);
stream_fn pick_stream_cta_ns4(int nd, int nlin);
stream_fn pick_stream_cta_ns5(int nd, int nlin);
stream_fn pick_stream_cta_ns6(int nd, int nlin);
stream_fn pick_stream_cta_ns7(int nd, int nlin);
stream_fn pick_stream_cta_ns8(int nd, int nlin);

stream_fn pick_stream_cta_kernel(int ns, int ndeep, int nlin) {
  switch (ns) {
    case 1: return pick_stream_cta_ns1(ndeep, nlin);
    case 2: return pick_stream_cta_ns2(ndeep, nlin);
    case 3: return pick_stream_cta_ns3(ndeep, nlin);
    case 4: return pick_stream_cta_ns4(ndeep, nlin);
    case 5: return pick_stream_cta_ns5(ndeep, nlin);
    case 6: return pick_stream_cta_ns6(ndeep, nlin);
    case 7: return pick_stream_cta_ns7(ndeep, nlin);
    case 8: return pick_stream_cta_ns8(ndeep, nlin);
  }
  return nullptr;
}
stream_fn pick_stream_cta_max_kernel() { return (stream_fn)stream_cta_max_kernel<0>; }
unsigned stream_cta_smem_bytes(int nw, int ncol, int rows_total, int deep_entries) {
  const CtaLayout L = cta_layout(nw, ncol, rows_total, deep_entries);
  // 16-byte vector accesses everywhere (tables, staged blocks, accumulators): 0 = a layout bug, refused by the caller
  if ((L.ctl | L.tables | L.stages | L.stage_bytes | L.blk_bytes | L.deep) & 15u) return 0u;
  return L.total;
}
}  // namespace gwi
static_assert(gwi::CTA_STAGES == gwi::CTA_NSTAGE && gwi::CTA_WARPS_MAX == gwi::CTA_MAX_WARPS, "plan.cpp sizes the CTA geometry with the kernel's constants");
