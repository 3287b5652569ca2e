// placeholder until the tcgen05 kernel lands: every convolution uses the SIMT kernel
#include "common.cuh"
#include "model.cuh"
namespace lion {
int conv_tc_prepare(Model*, ConvW&) { return 0; }
int conv_tc_pack_job(const PackJob&) { return 0; }
bool conv_tc_usable(const ConvW&, const ConvGeom&) { return false; }
int conv_tc_run(Ctx*, const ConvW&, const float4*, int, float4*, int, double*, double*, const ConvGeom&, int) { return LION_ERR_STATE; }
}
