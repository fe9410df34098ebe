// Descriptor probe (test infrastructure for the kernel author).  Built into its OWN library
// (libcuriosity_b200_probe.so, `python -m curiosity_baselines_b200.build --probe`), never into the product .so:
// copies caller-supplied raw shared-memory images for A and B, issues `ksteps` tcgen05.mma with
// caller-supplied descriptor fields and reads the accumulator back.  Lets the tests pin down the
// hardware semantics of MN-major layouts and of the descriptor base_offset on B200 without
// recompiling.  Single CTA; never on the product path.
#include "common.cuh"
#include "umma_common.cuh"

namespace cb { thread_local char g_err[512] = ""; }     // this library's own error text (api.cu is not linked in)
extern "C" const char* cb_probe_last_error(void) { return cb::g_err; }
static int probe_device_ok() {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

namespace {

struct ProbeParams {
    const uint8_t* a_img; const uint8_t* b_img; int a_bytes, b_bytes;   // images copied to smem (1024-aligned bases)
    int a_start, a_lbo, a_sbo, a_base_off, a_kstep, a_layout;         // A descriptor fields (bytes; layout 0 = SW128)
    int b_start, b_lbo, b_sbo, b_base_off, b_kstep, b_layout;
    int a_mn_major, b_mn_major, M, N, ksteps;
    float* d_out;                                                     // [128][N] fp32 (TMEM lanes x columns)
};

__device__ __forceinline__ uint64_t probe_desc(uint32_t saddr, int lbo, int sbo, int base_off, int layout) {
    uint64_t d = umma::make_smem_desc(saddr, (uint32_t)lbo, (uint32_t)sbo);
    d |= (uint64_t)(base_off & 7) << 49;
    if (layout) d = (d & ~((uint64_t)7 << 61)) | ((uint64_t)(layout & 7) << 61);
    return d;
}

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - umma::smem_u32(smem_raw));
    uint8_t* a_s = gbase;
    uint8_t* b_s = gbase + ((p.a_bytes + 1023) & ~1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < p.a_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(p.a_img)[i];
    for (int i = threadIdx.x; i < p.b_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(p.b_img)[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { umma::mbar_init(umma::smem_u32(&bar), 1); umma::fence_barrier_init(); }
    if (warp == 0) umma::tmem_alloc(umma::smem_u32(&tmem_slot), 256);
    // generic-proxy smem writes must be visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma::make_idesc(p.M, p.N, p.a_mn_major, p.b_mn_major);
        const uint32_t sa = umma::smem_u32(a_s), sb = umma::smem_u32(b_s);
        for (int k = 0; k < p.ksteps; ++k) {
            uint64_t da = probe_desc(sa + p.a_start + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_base_off, p.a_layout);
            uint64_t db = probe_desc(sb + p.b_start + k * p.b_kstep, p.b_lbo, p.b_sbo, p.b_base_off, p.b_layout);
            umma::mma_f16_ss(tmem, da, db, idesc, k ? 1u : 0u);
        }
        umma::mma_commit(umma::smem_u32(&bar));
    }
    umma::mbar_wait(umma::smem_u32(&bar), 0);
    umma::fence_after_thread_sync();
    for (int c0 = 0; c0 < p.N; c0 += 32) {      // reads whole 32-column groups; columns >= N are not stored
        uint32_t raw[32];
        umma::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, raw);
        umma::tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < p.N; ++j)
            p.d_out[(size_t)(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(raw[j]);
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) { umma::fence_after_thread_sync(); umma::tmem_dealloc(tmem, 256); }
}

}  // namespace

extern "C" int cb_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, const int* a_desc,
                             const int* b_desc, int a_mn_major, int b_mn_major, int M, int N, int ksteps,
                             float* d_out, void* stream) {
    CB_REQUIRE(a_img && b_img && a_desc && b_desc && d_out, CB_E_ARG, "cb_umma_probe: null pointer");
    CB_REQUIRE(probe_device_ok(), CB_E_ARCH, "cb_umma_probe: needs sm_100");
    CB_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes <= 200 * 1024 && N % 8 == 0 && N <= 256,
               CB_E_ARG, "cb_umma_probe: bad sizes");
    ProbeParams p{};
    p.a_img = (const uint8_t*)a_img; p.b_img = (const uint8_t*)b_img; p.a_bytes = a_bytes; p.b_bytes = b_bytes;
    p.a_start = a_desc[0]; p.a_lbo = a_desc[1]; p.a_sbo = a_desc[2]; p.a_base_off = a_desc[3]; p.a_kstep = a_desc[4]; p.a_layout = a_desc[5];
    p.b_start = b_desc[0]; p.b_lbo = b_desc[1]; p.b_sbo = b_desc[2]; p.b_base_off = b_desc[3]; p.b_kstep = b_desc[4]; p.b_layout = b_desc[5];
    p.a_mn_major = a_mn_major; p.b_mn_major = b_mn_major; p.M = M; p.N = N; p.ksteps = ksteps; p.d_out = d_out;
    int smem = a_bytes + b_bytes + 4096;
    CB_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<<<1, 128, smem, cb::as_stream(stream)>>>(p);
    CB_LAUNCH_CHECK();
    return CB_OK;
}
