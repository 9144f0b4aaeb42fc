// qk_sweep.cu -- fused sweep path (placeholder until the tuned kernels land: reports "not handled").
#include "qk_level.h"

int qk_fused_stage(qk_level *, const qk_hydro_params *, int, const qk_array4 *, const qk_array4 *, const qk_array4 *, double, int64_t *, cudaStream_t,
		   bool *handled)
{
	*handled = false;
	return 0;
}
