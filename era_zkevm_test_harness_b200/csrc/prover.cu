// placeholder, replaced by the real prover next
#include "host_common.cuh"
namespace zk { extern thread_local std::string g_last_error; }
extern "C" {
int zkgpu_setup_create(zkgpu_ctx*, const zkgpu_geometry*, const zkgpu_proof_config*, const uint64_t*, zkgpu_setup**, uint64_t*) { zk::g_last_error = "not implemented"; return 98; }
void zkgpu_setup_destroy(zkgpu_setup*) {}
int zkgpu_prove(zkgpu_ctx*, const zkgpu_setup*, const uint64_t*, uint64_t*, size_t) { zk::g_last_error = "not implemented"; return 98; }
int zkgpu_prove_device(zkgpu_ctx*, const zkgpu_setup*, const uint64_t*, uint64_t*, size_t) { zk::g_last_error = "not implemented"; return 98; }
}
