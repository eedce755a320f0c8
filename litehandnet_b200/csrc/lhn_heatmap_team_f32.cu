// Instantiation of the persistent team kernel for float heatmaps.
#define LHN_TEAM_DTYPE_TU
#define LHN_TRACE_EXPORT
#include "lhn_heatmap_team.cuh"

namespace lhn {
template int dispatch_team<float>(HmArgs&, bool, bool, int, int, size_t, cudaStream_t);
}
