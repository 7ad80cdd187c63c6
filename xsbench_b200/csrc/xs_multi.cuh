// xs_multi.cuh -- the only collective of the path: an NCCL all-reduce (sum, u64 x 2) of
// {verification, n_lookups} across the GPUs of one node, used when one process drives
// several devices (xs_gpu_init with n_gpus > 1).  One-process-per-GPU callers (bench.py under
// torchrun) shard with xs_gpu_run_range and all-reduce through torch.distributed instead.
//
// The reference has no multi-GPU decomposition at all (its MPI mode replicates the whole
// problem per rank and sums lookups/s once: openmp-threading/io.c:51-56); lookups are
// independent (cuda/Simulation.cu:53-56), so partition + all-reduce is exact.
//
// NCCL is loaded at run time (dlopen) so that single-GPU users carry no NCCL dependency.
// If it cannot be loaded the call fails with XS_ERR_NCCL -- there is no silent fallback.
#pragma once

#include <dlfcn.h>

namespace {

typedef struct ncclComm *xs_ncclComm_t;
typedef int xs_ncclResult_t;                  // ncclSuccess == 0
enum { XS_NCCL_UINT8 = 1, XS_NCCL_UINT64 = 5, XS_NCCL_SUM = 0 }; // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct MultiState {
    void *lib = nullptr;
    xs_ncclComm_t comm[8] = {};
    int n = 0;
    xs_ncclResult_t (*CommInitAll)(xs_ncclComm_t *, int, const int *) = nullptr;
    xs_ncclResult_t (*CommDestroy)(xs_ncclComm_t) = nullptr;
    xs_ncclResult_t (*GroupStart)() = nullptr;
    xs_ncclResult_t (*GroupEnd)() = nullptr;
    xs_ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, xs_ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(xs_ncclResult_t) = nullptr;
};

int xs_multi_allreduce(xs_gpu_ctx *ctx);

int xs_multi_init(xs_gpu_ctx *ctx)
{
    MultiState *m = new MultiState;
    ctx->nccl = m;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) if (!m->lib) m->lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!m->lib) return set_error(XS_ERR_NCCL, "cannot load NCCL: %s", dlerror());
    *(void **)&m->CommInitAll = dlsym(m->lib, "ncclCommInitAll");
    *(void **)&m->CommDestroy = dlsym(m->lib, "ncclCommDestroy");
    *(void **)&m->GroupStart = dlsym(m->lib, "ncclGroupStart");
    *(void **)&m->GroupEnd = dlsym(m->lib, "ncclGroupEnd");
    *(void **)&m->AllReduce = dlsym(m->lib, "ncclAllReduce");
    *(void **)&m->GetErrorString = dlsym(m->lib, "ncclGetErrorString");
    if (!m->CommInitAll || !m->CommDestroy || !m->GroupStart || !m->GroupEnd || !m->AllReduce)
        return set_error(XS_ERR_NCCL, "NCCL library lacks required symbols");
    int devs[8];
    m->n = (int)ctx->dev.size();
    for (int g = 0; g < m->n; g++) devs[g] = ctx->dev[g].device;
    xs_ncclResult_t r = m->CommInitAll(m->comm, m->n, devs);
    if (r != 0) {
        m->n = 0;
        return set_error(XS_ERR_NCCL, "ncclCommInitAll failed: %s", m->GetErrorString ? m->GetErrorString(r) : "?");
    }
    // the first collective sets up channels and buffers (~60 ms): do it here, not in the first run
    int rc = xs_multi_allreduce(ctx);
    for (int g = 0; g < m->n && rc == XS_OK; g++) {
        DeviceState &d = ctx->dev[g];
        if (cudaSetDevice(d.device) != cudaSuccess || cudaStreamSynchronize(d.stream) != cudaSuccess)
            rc = set_error(XS_ERR_CUDA, "NCCL warm-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return rc;
}

int xs_multi_allreduce(xs_gpu_ctx *ctx)
{
    MultiState *m = static_cast<MultiState *>(ctx->nccl);
    if (!m || m->n != (int)ctx->dev.size()) return set_error(XS_ERR_NCCL, "NCCL communicators not initialised");
    xs_ncclResult_t r = m->GroupStart();
    for (int g = 0; g < m->n && r == 0; g++) {
        DeviceState &d = ctx->dev[g];
        r = m->AllReduce(d.accum, d.accum, 2, XS_NCCL_UINT64, XS_NCCL_SUM, m->comm[g], d.stream);
    }
    xs_ncclResult_t r2 = m->GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) return set_error(XS_ERR_NCCL, "ncclAllReduce failed: %s", m->GetErrorString ? m->GetErrorString(r) : "?");
    return XS_OK;
}

// History mode on an energy-band-sharded grid: the particles' feedback bytes (n_forward, written by the
// band that performed the lookup, zero elsewhere) summed over the devices.
int xs_multi_allreduce_bytes(xs_gpu_ctx *ctx, size_t n_bytes)
{
    MultiState *m = static_cast<MultiState *>(ctx->nccl);
    if (!m || m->n != (int)ctx->dev.size()) return set_error(XS_ERR_NCCL, "NCCL communicators not initialised");
    xs_ncclResult_t r = m->GroupStart();
    for (int g = 0; g < m->n && r == 0; g++) {
        DeviceState &d = ctx->dev[g];
        r = m->AllReduce(d.hist_fwd, d.hist_fwd, n_bytes, XS_NCCL_UINT8, XS_NCCL_SUM, m->comm[g], d.stream);
    }
    xs_ncclResult_t r2 = m->GroupEnd();
    if (r == 0) r = r2;
    if (r != 0) return set_error(XS_ERR_NCCL, "ncclAllReduce (history feedback) failed: %s", m->GetErrorString ? m->GetErrorString(r) : "?");
    return XS_OK;
}

void xs_multi_destroy(xs_gpu_ctx *ctx)
{
    MultiState *m = static_cast<MultiState *>(ctx->nccl);
    if (!m) return;
    for (int g = 0; g < m->n; g++) if (m->comm[g] && m->CommDestroy) m->CommDestroy(m->comm[g]);
    // the library stays loaded (NCCL keeps background state); only our handles are dropped
    delete m;
    ctx->nccl = nullptr;
}

}  // namespace
