// growbuf.h -- device buffers that GROW IN PLACE (CUDA virtual memory management).
//
// A shard's slabs (rows, fp16 shadow plane, norms, labels, tombstone bitmap) are append-only: searches read the
// rows below an atomically published count, writers fill the rows above it.  hnswlib.Index.resize_index -- and
// the reference's "rebuild when full" (src/datanode/handler.py:237-251) -- would otherwise mean allocate + copy +
// free under an exclusive lock (2 GB per million rows).  Here every slab is a reserved VIRTUAL address range
// (cuMemAddressReserve) that physical memory is mapped into as the shard grows (cuMemCreate + cuMemMap): the base
// pointer never changes, nothing is copied, and searches keep running while the tail is being mapped.
// Only when the reservation itself is exhausted are the same physical chunks re-mapped into a larger range
// (still no copy); that step needs the shard quiesced and is the caller's to serialise.
//
// The driver entry points are resolved at run time (cudaGetDriverEntryPoint): the library does not link libcuda,
// so it still loads -- and fails loudly on the first compute call -- on a box without a driver.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

namespace vdbk {

struct VmmApi {
    CUresult (*GetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    bool ok = false;
};

inline const VmmApi& vmm_api() {
    static VmmApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        bool ok = true;
        ok &= get("cuMemGetAllocationGranularity", (void**)&api.GetAllocationGranularity);
        ok &= get("cuMemAddressReserve", (void**)&api.AddressReserve);
        ok &= get("cuMemAddressFree", (void**)&api.AddressFree);
        ok &= get("cuMemCreate", (void**)&api.Create);
        ok &= get("cuMemRelease", (void**)&api.Release);
        ok &= get("cuMemMap", (void**)&api.Map);
        ok &= get("cuMemUnmap", (void**)&api.Unmap);
        ok &= get("cuMemSetAccess", (void**)&api.SetAccess);
        if (!ok) cudaGetLastError();
        api.ok = ok;
    });
    return api;
}

class GrowBuf {
public:
    GrowBuf() = default;
    GrowBuf(const GrowBuf&) = delete;
    GrowBuf& operator=(const GrowBuf&) = delete;
    ~GrowBuf() { release(); }

    void* ptr() const { return reinterpret_cast<void*>(va_); }
    size_t mapped() const { return mapped_; }
    size_t reserved() const { return va_bytes_; }
    size_t chunks() const { return chunks_.size(); }

    // Reserve `va_bytes` of address space on `device` (no physical memory yet).  The current context must be
    // the device's (cudaSetDevice).
    bool reserve(int device, size_t va_bytes, std::string& err) {
        const VmmApi& a = vmm_api();
        if (!a.ok) { err = "CUDA virtual memory management entry points are unavailable"; return false; }
        release();
        device_ = device;
        prop_ = CUmemAllocationProp{};
        prop_.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop_.location.id = device;
        if (!check(a.GetAllocationGranularity(&gran_, &prop_, CU_MEM_ALLOC_GRANULARITY_MINIMUM), "cuMemGetAllocationGranularity", err)) return false;
        va_bytes_ = round_up(va_bytes ? va_bytes : 1);
        if (!check(a.AddressReserve(&va_, va_bytes_, 0, 0, 0), "cuMemAddressReserve", err)) { va_ = 0; va_bytes_ = 0; return false; }
        return true;
    }

    // Back [0, bytes) with physical memory (maps one more chunk when the mapped prefix is shorter).  Existing
    // contents and the base pointer are untouched; safe while kernels read the already-mapped prefix.
    // false + err: out of memory, or bytes beyond the reservation (check reserved() first and call rebase()).
    bool ensure(size_t bytes, std::string& err) {
        const VmmApi& a = vmm_api();
        if (bytes <= mapped_) return true;
        const size_t want = round_up(bytes);
        if (want > va_bytes_) { err = "GrowBuf: reservation exhausted"; return false; }
        const size_t add = want - mapped_;
        CUmemGenericAllocationHandle h;
        if (!check(a.Create(&h, add, &prop_, 0), "cuMemCreate", err)) return false;
        if (!check(a.Map(va_ + mapped_, add, 0, h, 0), "cuMemMap", err)) { a.Release(h); return false; }
        CUmemAccessDesc acc{};
        acc.location = prop_.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (!check(a.SetAccess(va_ + mapped_, add, &acc, 1), "cuMemSetAccess", err)) {
            a.Unmap(va_ + mapped_, add);
            a.Release(h);
            return false;
        }
        chunks_.push_back({h, add});
        mapped_ = want;
        return true;
    }

    // Move to a larger reservation: the SAME physical chunks are mapped at the same offsets of a new range (no
    // copy), the old range is unmapped and freed.  The base pointer changes: nothing may be using the buffer.
    bool rebase(size_t new_va_bytes, std::string& err) {
        const VmmApi& a = vmm_api();
        const size_t nb = round_up(new_va_bytes);
        if (nb <= va_bytes_) return true;
        CUdeviceptr nva = 0;
        if (!check(a.AddressReserve(&nva, nb, 0, 0, 0), "cuMemAddressReserve", err)) return false;
        size_t off = 0;
        bool ok = true;
        for (const Chunk& c : chunks_) {
            ok = check(a.Map(nva + off, c.bytes, 0, c.h, 0), "cuMemMap", err);
            if (!ok) break;
            off += c.bytes;
        }
        if (ok && off) {
            CUmemAccessDesc acc{};
            acc.location = prop_.location;
            acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            ok = check(a.SetAccess(nva, off, &acc, 1), "cuMemSetAccess", err);
        }
        if (!ok) {
            if (off) a.Unmap(nva, off);
            a.AddressFree(nva, nb);
            return false;
        }
        if (mapped_) a.Unmap(va_, mapped_);
        a.AddressFree(va_, va_bytes_);
        va_ = nva;
        va_bytes_ = nb;
        return true;
    }

    void release() {
        if (!va_) return;
        const VmmApi& a = vmm_api();
        if (mapped_) a.Unmap(va_, mapped_);
        for (const Chunk& c : chunks_) a.Release(c.h);
        a.AddressFree(va_, va_bytes_);
        chunks_.clear();
        va_ = 0; va_bytes_ = 0; mapped_ = 0;
    }

private:
    struct Chunk { CUmemGenericAllocationHandle h; size_t bytes; };
    size_t round_up(size_t b) const { return (b + gran_ - 1) / gran_ * gran_; }
    static bool check(CUresult r, const char* what, std::string& err) {
        if (r == CUDA_SUCCESS) return true;
        err = std::string(what) + " failed (CUresult " + std::to_string((int)r) + (r == CUDA_ERROR_OUT_OF_MEMORY ? ", out of memory)" : ")");
        return false;
    }
    CUdeviceptr va_ = 0;
    size_t va_bytes_ = 0, mapped_ = 0, gran_ = 2u << 20;
    int device_ = 0;
    CUmemAllocationProp prop_{};
    std::vector<Chunk> chunks_;
};

}  // namespace vdbk
