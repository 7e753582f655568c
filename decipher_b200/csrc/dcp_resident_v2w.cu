// dcp_resident_v2w.cu -- the WIDE build of the second-generation resident kernel (dcp_resident_v2.cu): one team of
// 8 compute + 4 river warps per CTA, 1024 unit columns per member slot, for catchments of 513 .. 1024 HRUs
// (k_physics_resident_v2w, dcp_launch_physics_resident_v2w).  Same source, different tile constants and table layout.
#define R2_WIDE 1
#include "dcp_resident_v2.cu"
