// flat.cu — one flat virtual address range over the shards of every GPU of the node.
//
// A sharded state keeps the top g index bits in the rank number.  With the CUDA virtual
// memory management API every rank allocates its shard as an exportable physical
// allocation, imports the other ranks' allocations (POSIX file descriptors, passed between
// the processes by the host layer) and maps all R of them back to back into ONE reserved
// virtual range: amplitude k of the WHOLE 2^n state lives at base + 8k on every GPU, local
// or over NVLink.  The fused tile kernel then runs unchanged on the whole state — a pass
// whose tile contains a rank bit simply loads part of its tile from peer memory and stores
// it back there, overlapping the NVLink traffic with its own arithmetic tile by tile — and
// every rank processes 1/R of the tiles of every pass (plan.cu, aqs_plan_run_shard).  No
// qubit remap, no staging buffers, no collective on the data path.
//
// The driver entry points are fetched with cudaGetDriverEntryPoint, so the library has no
// link-time dependency on libcuda (it must load on machines without a driver).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstring>
#include <new>
#include <vector>

#include "engine_internal.h"

struct aqs_flat_view_s {
    CUdeviceptr base = 0;
    size_t bytes = 0;
    std::vector<std::pair<CUdeviceptr, size_t>> maps;    // mapped sub-ranges
};

struct aqs_flat_s {
    int world = 0, rank = 0, device = 0;
    size_t shard_bytes = 0;
    CUdeviceptr base = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    std::vector<char> mapped;
    int own_fd = -1;
    // staged passes: local staging memory for the peers' parts of a pass's tiles, and views of the state in which
    // those parts ARE the staging memory
    // (cuMemMap cannot map a PART of an allocation — its offset must be zero — so the staging memory is a pool of
    // block-sized allocations, shared by all views: only one staged pass runs at a time)
    std::vector<std::pair<size_t, CUmemGenericAllocationHandle>> stage_blocks;    // (bytes, handle)
    size_t granularity = 0;
    std::vector<aqs_flat_view_s*> views;
};

namespace {

struct Drv {
    decltype(&cuMemCreate) MemCreate = nullptr;
    decltype(&cuMemRelease) MemRelease = nullptr;
    decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
    decltype(&cuMemAddressFree) MemAddressFree = nullptr;
    decltype(&cuMemMap) MemMap = nullptr;
    decltype(&cuMemUnmap) MemUnmap = nullptr;
    decltype(&cuMemSetAccess) MemSetAccess = nullptr;
    decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
    decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
    decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
Drv g_drv;

template <typename F>
bool entry(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

bool load_driver() {
    if (g_drv.ok) return true;
    Drv d;
    bool ok = entry("cuMemCreate", d.MemCreate) && entry("cuMemRelease", d.MemRelease) &&
              entry("cuMemAddressReserve", d.MemAddressReserve) && entry("cuMemAddressFree", d.MemAddressFree) &&
              entry("cuMemMap", d.MemMap) && entry("cuMemUnmap", d.MemUnmap) && entry("cuMemSetAccess", d.MemSetAccess) &&
              entry("cuMemExportToShareableHandle", d.MemExportToShareableHandle) &&
              entry("cuMemImportFromShareableHandle", d.MemImportFromShareableHandle) &&
              entry("cuMemGetAllocationGranularity", d.MemGetAllocationGranularity) && entry("cuGetErrorString", d.GetErrorString);
    if (!ok) return false;
    d.ok = true;
    g_drv = d;
    return true;
}

int fail_drv(CUresult r, const char* what) {
    const char* msg = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &msg);
    char buf[384];
    snprintf(buf, sizeof buf, "CUDA driver error %d (%s): %s", (int)r, msg ? msg : "?", what);
    return aqs::fail(AQS_ERR_CUDA, buf);
}
#define DRV_TRY(x)                                      \
    do {                                                \
        CUresult r_ = (x);                              \
        if (r_ != CUDA_SUCCESS) return fail_drv(r_, #x); \
    } while (0)

CUmemAllocationProp shard_prop(int device) {
    CUmemAllocationProp prop;
    std::memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

int map_slot(aqs_flat_s* f, int slot, CUmemGenericAllocationHandle h) {
    DRV_TRY(g_drv.MemMap(f->base + (CUdeviceptr)slot * f->shard_bytes, f->shard_bytes, 0, h, 0));
    f->handles[slot] = h;
    f->mapped[slot] = 1;
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = f->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DRV_TRY(g_drv.MemSetAccess(f->base + (CUdeviceptr)slot * f->shard_bytes, f->shard_bytes, &acc, 1));
    return AQS_OK;
}

}  // namespace

extern "C" {

int aqs_flat_create(uint64_t shard_bytes, int world, int rank, aqs_flat_t* out, int* fd_out) {
    if (!out || !fd_out) return aqs::fail(AQS_ERR_INVALID, "null argument");
    if (world < 1 || rank < 0 || rank >= world || shard_bytes == 0) return aqs::fail(AQS_ERR_INVALID, "bad shard geometry");
    int device = 0;
    if (aqs_engine_device(&device, nullptr, nullptr) != AQS_OK) return AQS_ERR_STATE;
    if (!load_driver()) return aqs::fail(AQS_ERR_STATE, "CUDA virtual memory management entry points are not available");
    CUmemAllocationProp prop = shard_prop(device);
    size_t gran = 0;
    DRV_TRY(g_drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    if (shard_bytes % gran) {
        char buf[160];
        snprintf(buf, sizeof buf, "shard of %llu bytes is not a multiple of the %zu-byte mapping granularity", (unsigned long long)shard_bytes, gran);
        return aqs::fail(AQS_ERR_INVALID, buf);
    }
    aqs_flat_s* f = new (std::nothrow) aqs_flat_s();
    if (!f) return aqs::fail(AQS_ERR_NOMEM, "host allocation failed");
    f->world = world; f->rank = rank; f->device = device; f->shard_bytes = (size_t)shard_bytes;
    f->granularity = gran;
    f->handles.assign(world, 0);
    f->mapped.assign(world, 0);
    CUmemGenericAllocationHandle h = 0;
    CUresult r = g_drv.MemCreate(&h, f->shard_bytes, &prop, 0);
    if (r == CUDA_ERROR_OUT_OF_MEMORY) {
        // cuMemCreate does not see the engine's cache of recycled state buffers: give them back and retry once
        aqs_pool_trim();
        r = g_drv.MemCreate(&h, f->shard_bytes, &prop, 0);
    }
    if (r != CUDA_SUCCESS) { delete f; return r == CUDA_ERROR_OUT_OF_MEMORY ? aqs::fail(AQS_ERR_NOMEM, "cuMemCreate: out of device memory") : fail_drv(r, "cuMemCreate"); }
    r = g_drv.MemAddressReserve(&f->base, f->shard_bytes * (size_t)world, f->shard_bytes < (1ull << 21) ? 0 : (1ull << 21), 0, 0);
    if (r != CUDA_SUCCESS) { g_drv.MemRelease(h); delete f; return fail_drv(r, "cuMemAddressReserve"); }
    int rc = map_slot(f, rank, h);
    if (rc != AQS_OK && !f->handles[rank]) g_drv.MemRelease(h);      // cuMemMap failed before the handle was recorded
    if (rc == AQS_OK) {
        r = g_drv.MemExportToShareableHandle(&f->own_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) rc = fail_drv(r, "cuMemExportToShareableHandle");
    }
    if (rc != AQS_OK) { aqs_flat_destroy(f); return rc; }
    *fd_out = f->own_fd;
    *out = f;
    return AQS_OK;
}

int aqs_flat_attach(aqs_flat_t f, int peer_rank, int fd) {
    if (!f) return aqs::fail(AQS_ERR_INVALID, "null handle");
    if (peer_rank < 0 || peer_rank >= f->world || peer_rank == f->rank || f->mapped[peer_rank]) return aqs::fail(AQS_ERR_INVALID, "bad peer rank");
    CUmemGenericAllocationHandle h = 0;
    DRV_TRY(g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    int rc = map_slot(f, peer_rank, h);
    if (rc != AQS_OK && !f->handles[peer_rank]) g_drv.MemRelease(h);   // (a recorded handle is released by aqs_flat_destroy)
    return rc;
}

int aqs_flat_ptr(aqs_flat_t f, void** base, void** own_shard) {
    if (!f) return aqs::fail(AQS_ERR_INVALID, "null handle");
    if (base) *base = (void*)f->base;
    if (own_shard) *own_shard = (void*)(f->base + (CUdeviceptr)f->rank * f->shard_bytes);
    return AQS_OK;
}

// ---- staged passes ---------------------------------------------------------------------------------------------
// A pass whose tiles contain rank bits needs, on every GPU, parts of the peers' shards.  Reading them from inside the
// kernel means scattered 256-byte reads over NVLink, which run at about half the link rate (measured on 2 x B200: 425 GB/s
// against 790 GB/s for contiguous reads; peer WRITES reach 718 GB/s whatever the pattern: profiles/r02_peer_bw.txt).  The
// parts a rank needs are large contiguous blocks though (the tiles of a rank are selected by pinning high local bits), so
// the copy engines fetch them into local STAGING memory in big pieces while the kernel computes the previous chunk of
// tiles, and the kernel reads a VIEW of the state — a second virtual range in which this rank's shard is mapped as usual
// and the needed blocks of the peers' slots are backed by local staging allocations (the copies write through the view's
// own addresses) — while it still writes its results directly into the peers' HBM.
static void view_free(aqs_flat_view_s* v) {
    for (auto& m : v->maps) g_drv.MemUnmap(m.first, m.second);
    if (v->base) g_drv.MemAddressFree(v->base, v->bytes);
    delete v;
}

int aqs_flat_view_create(aqs_flat_t f, const aqs_flat_block* blocks, uint64_t n_blocks, void** view_base) {
    if (!f || !view_base || (!blocks && n_blocks)) return aqs::fail(AQS_ERR_INVALID, "null argument");
    const size_t total = f->shard_bytes * (size_t)f->world;
    for (uint64_t i = 0; i < n_blocks; ++i) {
        const aqs_flat_block& b = blocks[i];
        if (b.bytes == 0 || b.bytes % f->granularity || b.state_offset % f->granularity || b.state_offset + b.bytes > total)
            return aqs::fail(AQS_ERR_INVALID, "staged block is misaligned or out of range");
        const size_t s0 = b.state_offset / f->shard_bytes, s1 = (b.state_offset + b.bytes - 1) / f->shard_bytes;
        if (s0 != s1 || (int)s0 == f->rank) return aqs::fail(AQS_ERR_INVALID, "a staged block must lie inside ONE peer shard");
    }
    aqs_flat_view_s* v = new (std::nothrow) aqs_flat_view_s();
    if (!v) return aqs::fail(AQS_ERR_NOMEM, "host allocation failed");
    v->bytes = total;
    CUresult r = g_drv.MemAddressReserve(&v->base, total, 1ull << 21, 0, 0);
    if (r != CUDA_SUCCESS) { delete v; return fail_drv(r, "cuMemAddressReserve(view)"); }
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = f->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    auto map = [&](CUdeviceptr va, size_t bytes, CUmemGenericAllocationHandle h) {
        CUresult q = g_drv.MemMap(va, bytes, 0, h, 0);
        if (q != CUDA_SUCCESS) return q;
        v->maps.push_back({va, bytes});
        return g_drv.MemSetAccess(va, bytes, &acc, 1);
    };
    r = map(v->base + (CUdeviceptr)f->rank * f->shard_bytes, f->shard_bytes, f->handles[f->rank]);
    // staging blocks from the pool: the k-th block of a given size of this view takes the k-th pooled allocation of that size
    CUmemAllocationProp prop;
    std::memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = f->device;
    std::vector<char> used(f->stage_blocks.size(), 0);
    for (uint64_t i = 0; r == CUDA_SUCCESS && i < n_blocks; ++i) {
        size_t k = 0;
        while (k < f->stage_blocks.size() && (used[k] || f->stage_blocks[k].first != blocks[i].bytes)) ++k;
        if (k == f->stage_blocks.size()) {
            CUmemGenericAllocationHandle h = 0;
            r = g_drv.MemCreate(&h, blocks[i].bytes, &prop, 0);
            if (r == CUDA_ERROR_OUT_OF_MEMORY) {
                aqs_pool_trim();
                r = g_drv.MemCreate(&h, blocks[i].bytes, &prop, 0);
            }
            if (r != CUDA_SUCCESS) break;
            f->stage_blocks.push_back({(size_t)blocks[i].bytes, h});
            used.push_back(0);
        }
        used[k] = 1;
        r = map(v->base + blocks[i].state_offset, blocks[i].bytes, f->stage_blocks[k].second);
    }
    if (r != CUDA_SUCCESS) { view_free(v); return r == CUDA_ERROR_OUT_OF_MEMORY ? aqs::fail(AQS_ERR_NOMEM, "staging memory: out of device memory") : fail_drv(r, "mapping a view of the state"); }
    f->views.push_back(v);
    *view_base = (void*)v->base;
    return AQS_OK;
}

// plain asynchronous device-to-device copy on a stream (local or peer addresses of the flat range, the staging buffer)
int aqs_memcpy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
    if (!dst || !src) return aqs::fail(AQS_ERR_INVALID, "null argument");
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return aqs::fail_cuda(e, "cudaMemcpyAsync(device to device)", __LINE__);
    return AQS_OK;
}

int aqs_flat_destroy(aqs_flat_t f) {
    if (!f) return AQS_OK;
    cudaDeviceSynchronize();
    for (aqs_flat_view_s* v : f->views) view_free(v);
    f->views.clear();
    for (auto& sb : f->stage_blocks) g_drv.MemRelease(sb.second);
    f->stage_blocks.clear();
    for (int s = 0; s < f->world; ++s) {
        if (f->mapped[s]) g_drv.MemUnmap(f->base + (CUdeviceptr)s * f->shard_bytes, f->shard_bytes);
        if (f->handles[s]) g_drv.MemRelease(f->handles[s]);
    }
    if (f->base) g_drv.MemAddressFree(f->base, f->shard_bytes * (size_t)f->world);
    if (f->own_fd >= 0) close(f->own_fd);
    delete f;
    return AQS_OK;
}

}  // extern "C"
