// placeholder until the tcgen05 kernel lands
#include "common.h"
namespace mfa {
bool fwd_tc_eligible(const AttnParams&) { return false; }
cudaError_t launch_fwd_tc(const AttnParams&, cudaStream_t) { return cudaErrorNotSupported; }
}
