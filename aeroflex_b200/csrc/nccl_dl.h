// nccl_dl.h -- NCCL resolved at run time (dlopen) instead of at link time.
//
// The host process may already carry an NCCL (PyTorch bundles its own
// libnccl.so.2); linking a second copy by name makes whichever loads first win
// for everybody.  Resolving lazily means: if the process already has an NCCL we
// use that one, otherwise the system library, and single-GPU users never touch
// NCCL at all.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <stdexcept>
#include <string>

namespace afx {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;

    static NcclApi& get()
    {
        static NcclApi api = load();
        return api;
    }

private:
    template <class F>
    static void sym(void* h, const char* name, F& f)
    {
        f = reinterpret_cast<F>(dlsym(h, name));
        if (!f) throw std::runtime_error(std::string("NCCL symbol not found: ") + name);
    }
    static NcclApi load()
    {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // an NCCL the process already has
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) throw std::runtime_error(std::string("cannot load NCCL: ") + dlerror());
        NcclApi a;
        sym(h, "ncclGetUniqueId", a.GetUniqueId);
        sym(h, "ncclCommInitRank", a.CommInitRank);
        sym(h, "ncclCommDestroy", a.CommDestroy);
        sym(h, "ncclSend", a.Send);
        sym(h, "ncclRecv", a.Recv);
        sym(h, "ncclAllReduce", a.AllReduce);
        sym(h, "ncclGroupStart", a.GroupStart);
        sym(h, "ncclGroupEnd", a.GroupEnd);
        sym(h, "ncclGetErrorString", a.GetErrorString);
        sym(h, "ncclGetVersion", a.GetVersion);
        return a;
    }
};

}  // namespace afx
