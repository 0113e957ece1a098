// gemm_topk.cu -- K2 batched tcgen05 search (placeholder until the kernel lands: reports
// "unsupported" so every batch takes the scan kernel).
#include "gemm_topk.h"
namespace vdbk {
bool gemm_topk_supported(int, int, bool, int, size_t) { return false; }
cudaError_t gemm_topk_search(GemmPlan&, GemmWorkspace&, const GemmSearchArgs&, cudaStream_t, std::string& err) {
    err = "tensor path not built";
    return cudaErrorNotSupported;
}
void gemm_plan_free(GemmPlan&) {}
void gemm_workspace_free(GemmWorkspace&) {}
long gemm_plan_fallbacks(const GemmPlan&) { return 0; }
}  // namespace vdbk
