// Host-side launcher of the persistent team kernel (lhn_heatmap_team.cuh); the kernels themselves are
// instantiated per dtype in lhn_heatmap_team_{f32,bf16,f16}.cu so the three compile in parallel.
#include "lhn_heatmap_team.cuh"
