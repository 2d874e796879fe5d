// emu_stage_ptx.h (HOST EMULATION) -- TEST INFRASTRUCTURE ONLY.  build_emu.py puts this file in place of the PTX
// primitives of rans_stage.cuh (mbarrier, cp.async.bulk, cp.async, ld/st.shared by shared-window address).  The copies
// are deferred to the matching wait (emu_runtime.cpp) and every access checks the alignment the instruction needs.
// Included inside namespace afx::<mode>.
__device__ __forceinline__ void emu_need(bool ok, const char* what)
{
    if (!ok) { fprintf(stderr, "afx_emu: %s (block %u thread %u)\n", what, blockIdx.x, threadIdx.x); abort(); }
}
// a shared-window address is the byte offset into the block's dynamic shared memory plus 1024 (never 0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const unsigned char*>(p) - ::afx_emu::g_dyn_smem) + 1024u; }
__device__ __forceinline__ unsigned char* emu_sp(uint32_t a, uint32_t bytes, uint32_t align)
{
    emu_need(a >= 1024u && (size_t)(a - 1024u) + bytes <= 232448u, "shared-memory access outside the block's allocation");
    emu_need(((a - 1024u) & (align - 1u)) == 0, "misaligned shared-memory access");
    return ::afx_emu::g_dyn_smem + (a - 1024u);
}
__device__ __forceinline__ unsigned char* emu_ld(uint32_t a, uint32_t bytes, uint32_t align) { unsigned char* p = emu_sp(a, bytes, align); ::afx_emu::smem_access(a - 1024u, bytes, false); return p; }
__device__ __forceinline__ unsigned char* emu_st(uint32_t a, uint32_t bytes, uint32_t align) { unsigned char* p = emu_sp(a, bytes, align); ::afx_emu::smem_access(a - 1024u, bytes, true); return p; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t) { emu_sp(bar, 8, 8); ::afx_emu::mbar_init(bar); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { ::afx_emu::mbar_expect(bar, bytes); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { ::afx_emu::mbar_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    emu_need((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0, "cp.async.bulk needs 16-byte aligned addresses and size");
    if (bytes) ::afx_emu::mbar_queue(bar, emu_sp(dst, bytes, 16), src, bytes);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    emu_need((reinterpret_cast<uintptr_t>(src) & 15u) == 0, "cp.async 16 needs a 16-byte aligned source");
    ::afx_emu::cpasync_queue(emu_sp(dst, 16, 16), src, 16);
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src)
{
    emu_need((reinterpret_cast<uintptr_t>(src) & 7u) == 0, "cp.async 8 needs an 8-byte aligned source");
    ::afx_emu::cpasync_queue(emu_sp(dst, 8, 8), src, 8);
}
__device__ __forceinline__ void cp_async_commit() { ::afx_emu::cpasync_commit(); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { ::afx_emu::cpasync_wait(N); }

__device__ __forceinline__ d4 lds_d4(uint32_t a) { d4 v; memcpy(&v, emu_ld(a, 32, 16), 32); return v; }
__device__ __forceinline__ void sts_d4(uint32_t a, const d4& v) { memcpy(emu_st(a, 32, 16), &v, 32); }
__device__ __forceinline__ double2 lds_d2(uint32_t a) { double2 v; memcpy(&v, emu_ld(a, 16, 16), 16); return v; }
__device__ __forceinline__ double lds_d(uint32_t a) { double v; memcpy(&v, emu_ld(a, 8, 8), 8); return v; }
__device__ __forceinline__ void sts_d(uint32_t a, double v) { memcpy(emu_st(a, 8, 8), &v, 8); }
__device__ __forceinline__ uint4 lds_u4(uint32_t a) { uint4 v; memcpy(&v, emu_ld(a, 16, 16), 16); return v; }
__device__ __forceinline__ void sts_u4(uint32_t a, const uint4& v) { memcpy(emu_st(a, 16, 16), &v, 16); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; memcpy(&v, emu_ld(a, 4, 4), 4); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { memcpy(emu_st(a, 4, 4), &v, 4); }

