// Instantiation of the persistent team kernel for __half heatmaps.
#define LHN_TEAM_DTYPE_TU
#include "lhn_heatmap_team.cuh"

namespace lhn {
template int dispatch_team<__half>(HmArgs&, bool, bool, int, int, size_t, cudaStream_t);
}
