// niq_tu_cp.cu -- launcher of the persistent closest_point kernel (niq_cp.cuh): windows of <= 2048 entries, fixed-row
// modes, nets whose weights stay resident in shared memory (hidden width <= 64).
#include "niq_internal.h"
#include "niq_cp.cuh"

template <int WMAX>
static int launch_cp_w(niq_ctx* c, NetDev net, int total_floats, const CpArgs& a, bool* fits) {
    using EB = Engine<WMAX, TileBox3>;
    using EP = Engine<WMAX, TilePts>;
    const size_t smem = std::max(place_weights<EB>(c, net, total_floats), EP::smem_bytes(total_floats + kResidentPad));
    *fits = net.resident != 0 && smem <= c->prop.sharedMemPerBlockOptin;
    if (!*fits) return NIQ_OK;
    auto kernel = k_cp_persistent<WMAX>;
    TRY(set_smem(kernel, smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
    if (per_sm < 1) { *fits = false; return NIQ_OK; }
    void* args[] = {(void*)&net, (void*)&a};
    LaunchTimer lt(c, 0);
    CU(cudaLaunchCooperativeKernel((const void*)kernel, dim3(c->prop.multiProcessorCount), dim3(kThreads), args, smem, c->stream));
    return NIQ_OK;
}

int launch_cp_persistent(niq_ctx* c, const niq_mlp* m, const CpArgs& a, bool* fits) {
    *fits = false;
    if (m->wmax == 32) return launch_cp_w<32>(c, m->net, m->total_floats, a, fits);
    if (m->wmax == 64) return launch_cp_w<64>(c, m->net, m->total_floats, a, fits);
    return NIQ_OK;
}
