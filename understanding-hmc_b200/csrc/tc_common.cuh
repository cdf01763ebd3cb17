// tcgen05 / TMEM building blocks shared by the tensor-core kernels (random_tc.cu, nuts_tc.cu), D = 100:
// tile constants, the bf16x3 / fp16x2 splits, tensor-memory loads and stores, the MMA sequence of one gradient pass.
#pragma once
#include "hmc_common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int TC_ND = 100;          // dimensions handled by this instantiation
constexpr int TC_KP = 112;          // padded K = N (multiple of 16)
constexpr int TC_KC = TC_KP / 8;    // 16-byte chunks (8 bf16) per operand row
constexpr int TC_M = 128;           // chains per CTA
constexpr int TC_SPL = 4;           // threads (dimension slices) per chain
constexpr int TC_THREADS = TC_M * TC_SPL;       // worker threads
constexpr int TC_NT = TC_THREADS + 128;         // + one warpgroup whose first warp only issues the MMAs (the issuing thread
                                                // is held while the tensor pipe drains; a worker in that role stalls the
                                                // whole CTA).  A whole warpgroup, so that it can hand its registers to the
                                                // workers (setmaxnreg): 4 x 112 + 24 registers x 32 lanes per scheduler (of the 5 x 96 x 32 allocated at launch).
constexpr int TC_BPART = TC_KC * TC_KP * 16;    // bytes of one B part
constexpr uint32_t TC_ACOL = 128;               // first TMEM column of the A parts (accumulator: columns 0..111)
constexpr uint32_t TC_APITCH = 64;              // TMEM columns per A part (56 used: K/2)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// named barriers: 1..4 = the four warps of a 32-chain group, 5 = workers + issuing warp.  The non-".aligned" PTX forms
// (barrier.sync / barrier.arrive count threads and do not require a converged warp): the worker loops are full of per-lane
// branches, and with the aligned forms the tensor-core NUTS kernel was observed to get a warp one barrier out of step
// (profiles/r2_nuts_tc_barrier_note.md).
__device__ __forceinline__ void bar_all() { asm volatile("barrier.sync 5, %0;" ::"n"(TC_NT) : "memory"); }
// the workers only ARRIVE at S1 (they never wait for each other there: what they need next is the MMA, through its mbarrier)
__device__ __forceinline__ void bar_all_arrive() { asm volatile("barrier.arrive 5, %0;" ::"n"(TC_NT) : "memory"); }
__device__ __forceinline__ void bar_group(int grp) { asm volatile("barrier.sync %0, 128;" ::"r"(1 + grp) : "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell); no swizzle, base offset 0
    return d;
}

// Split precisions of the gradient product (both FP32-grade, tests/test_bf16x3_split_cpu.py):
//   PREC_BF16X3 : x = b1 + b2 + b3 exactly (three bf16 parts), six part products  (1,3)(3,1)(2,2)(1,2)(2,1)(1,1)
//   PREC_FP16X2 : x = h1 + h2 + O(2^-23 x) (two fp16 parts), three part products  (1,2)(2,1)(1,1): the dropped terms are
//                 at the level of the float32 rounding of x itself; needs |x| < 6e4 (the host checks the start points).
// The residuals are formed with the mixed-precision subtract (one FHADD per value, half selector on the packed pair):
//   r = h1 - x = -(x - h1),  h2' = rn(r) = -h2,  s = h2' - r = x - h1 - h2,  h3 = rn(s)
// so the SECOND part comes out negated; the MMAs that read it set the negate-A bit of the instruction descriptor.
enum : int { PREC_BF16X3 = 0, PREC_FP16X2 = 1 };
template <int PREC> struct TcPrec;
template <> struct TcPrec<PREC_BF16X3> { static constexpr int NPART = 3, NPROD = 6; static constexpr bool F16 = false; };
template <> struct TcPrec<PREC_FP16X2> { static constexpr int NPART = 2, NPROD = 3; static constexpr bool F16 = true; };

template <int PREC>
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& h1, uint32_t& h2, uint32_t& h3) {
    float r0, r1;
    if constexpr (!TcPrec<PREC>::F16) {
        float s0, s1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(x1), "f"(x0));
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tsub.rn.f32.bf16 %0, lo, %3;\n\tsub.rn.f32.bf16 %1, hi, %4;\n\t}"
            : "=f"(r0), "=f"(r1) : "r"(h1), "f"(x0), "f"(x1));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(r1), "f"(r0));
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tsub.rn.f32.bf16 %0, lo, %3;\n\tsub.rn.f32.bf16 %1, hi, %4;\n\t}"
            : "=f"(s0), "=f"(s1) : "r"(h2), "f"(r0), "f"(r1));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h3) : "f"(s1), "f"(s0));
    } else {
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(x1), "f"(x0));
        asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tsub.rn.f32.f16 %0, lo, %3;\n\tsub.rn.f32.f16 %1, hi, %4;\n\t}"
            : "=f"(r0), "=f"(r1) : "r"(h1), "f"(x0), "f"(x1));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(r1), "f"(r0));
        h3 = 0u;
    }
}

// tensor-memory stores of one warp: lane i writes TMEM lane (32 * (warp % 4) + i), consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, const uint32_t* w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, const uint32_t* w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(w[0]), "r"(w[1]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// predicated in-place 128-bit shared-memory load
__device__ __forceinline__ void lds4_if(int pr, uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(a), "+f"(b), "+f"(c), "+f"(d) : "r"(addr), "r"(pr) : "memory");
}

// positions x[0..16) of a slice -> the parts, TMEM columns acol .. acol+7 of each part
template <int PREC>
__device__ __forceinline__ void put_half0(uint32_t acol, const float* x) {
    uint32_t w1[8], w2[8], w3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_pair<PREC>(x[2 * e], x[2 * e + 1], w1[e], w2[e], w3[e]);
    tmem_st8(acol, w1); tmem_st8(acol + TC_APITCH, w2);
    if constexpr (TcPrec<PREC>::NPART == 3) tmem_st8(acol + 2 * TC_APITCH, w3);
}
// positions x[16..24) (x[16..28) for the wide slice) -> columns acol+8 .. acol+11 (.. acol+13)
template <int PREC>
__device__ __forceinline__ void put_half1(uint32_t acol, const float* x, bool wide) {
    uint32_t w1[6], w2[6], w3[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) split_pair<PREC>(x[16 + 2 * e], x[17 + 2 * e], w1[e], w2[e], w3[e]);
    tmem_st4(acol + 8, w1); tmem_st4(acol + 8 + TC_APITCH, w2);
    if constexpr (TcPrec<PREC>::NPART == 3) tmem_st4(acol + 8 + 2 * TC_APITCH, w3);
    if (wide) {
        tmem_st2(acol + 12, w1 + 4); tmem_st2(acol + 12 + TC_APITCH, w2 + 4);
        if constexpr (TcPrec<PREC>::NPART == 3) tmem_st2(acol + 12 + 2 * TC_APITCH, w3 + 4);
    }
}

// The MMAs of a gradient pass: the part products, small terms first -- bf16x3: (1,3) (3,1) (2,2) (1,2) (2,1) (1,1), fp16x2:
// (1,2) (2,1) (1,1) -- seven K steps each.  A from tensor memory (part pa, 8 columns per K step; part 2 is stored negated,
// its products set the negate-A bit), B descriptor = constant high word + start address (>> 4) in the low word; the
// per-MMA offsets are immediates inside the asm so that nothing is hoisted into registers.
template <int PREC, int I>
__device__ __forceinline__ void tc_mma_all(uint32_t tmem, uint32_t dlo, uint32_t dhi, uint32_t idesc) {
    if constexpr (I < TcPrec<PREC>::NPROD * (TC_KP / 16)) {
        constexpr int pa3[6] = {0, 2, 1, 0, 1, 0}, pb3[6] = {2, 0, 1, 1, 0, 0};
        constexpr int pa2[3] = {0, 1, 0}, pb2[3] = {1, 0, 0};
        constexpr int t = I / (TC_KP / 16), ks = I % (TC_KP / 16);
        constexpr int pa = (PREC == PREC_BF16X3) ? pa3[t % 6] : pa2[t % 3], pb = (PREC == PREC_BF16X3) ? pb3[t % 6] : pb2[t % 3];
        asm volatile(
            "{\n\t.reg .pred pacc;\n\t.reg .b32 ta, bl, id;\n\t.reg .b64 db;\n\t"
            "setp.ne.b32 pacc, %4, 0;\n\t"
            "add.u32 ta, %0, %5;\n\t"
            "add.u32 bl, %1, %6;\n\t"
            "or.b32 id, %3, %8;\n\t"
            "mov.b64 db, {bl, %2};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, id, {%7, %7, %7, %7}, pacc;\n\t}"
            ::"r"(tmem), "r"(dlo), "r"(dhi), "r"(idesc), "r"(I ? 1u : 0u),
              "n"((int)(TC_ACOL + pa * TC_APITCH + 8 * ks)), "n"((pb * TC_BPART + ks * 2 * TC_KP * 16) >> 4), "r"(0u),
              "n"(pa == 1 ? (1 << 13) : 0));
        tc_mma_all<PREC, I + 1>(tmem, dlo, dhi, idesc);
    }
}


}  // namespace
