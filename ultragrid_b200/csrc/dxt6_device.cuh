// DXT5-YCoCg ("DXT6") 4x4 block encoder — device side.  See dxt_device.cuh for the contract: the operation
// tree below (which sub-expressions are double, where FMAs were contracted, which way products were
// grouped) was read from the SASS of UltraGrid's cuda_dxt/cuda_dxt.cu built by nvcc 12.9 for sm_100a and is
// restated with explicit-rounding intrinsics.  Reference source lines are cited per step.
#pragma once
#include "dxt_device.cuh"

namespace ugb {

__device__ constexpr float kOffset = 0.50196081399917602539f;  // (float)(128.0 / 255.0), cuda_dxt.cu:139

/// roundf(x) for x >= 0 followed by the u32 conversion, as ptxas emits it: trunc(add.rz(x, 0.5)).
/// Returned as the magic float 2^23 + q (floor through a round-down add, no F2I).
__device__ __forceinline__ float roundu_magic(float x)
{
        return __fadd_rd(__fadd_rz(x, 0.5f), kFloorMagic);
}
__device__ __forceinline__ uint32_t magic_bits(float m) { return __float_as_uint(m) - 0x4B000000u; }

/// Colour index of one pixel from its distances to the four palette entries (cuda_dxt.cu:337-345): b0 = d0 > d3, b1 = d1 > d2, b2 = d0 > d2,
/// b3 = d1 > d3, b4 = d2 > d3; index = (b0 & b4) | (((b1 & b2) | (b0 & b3)) << 1).  Written with predicate-combining compares
/// (FSETP...AND) and predicated ORs: 9 instructions where selects on 0/1 integers take 14.  `>` is false for NaN in both forms.
__device__ __forceinline__ void color_index_bits(uint32_t &cidx, float d0, float d1, float d2, float d3, uint32_t lo_bit, uint32_t hi_bit)
{
        asm("{\n\t"
            ".reg .pred q, s, p, r;\n\t"
            "setp.gt.f32 q, %1, %4;\n\t"         // b0
            "setp.gt.and.f32 s, %3, %4, q;\n\t"  // b4 & b0
            "setp.gt.f32 p, %2, %3;\n\t"         // b1
            "setp.gt.and.f32 p, %1, %3, p;\n\t"  // b2 & b1
            "setp.gt.and.f32 r, %2, %4, q;\n\t"  // b3 & b0
            "or.pred p, p, r;\n\t"
            "@s or.b32 %0, %0, %5;\n\t"
            "@p or.b32 %0, %0, %6;\n\t"
            "}"
            : "+r"(cidx)
            : "f"(d0), "f"(d1), "f"(d2), "f"(d3), "r"(lo_bit), "r"(hi_bit));
}

/// count = #{k : a <= T_k} of one pixel by binary search over the ordered thresholds (T0 >= T1 >= ... >= T6), its three bits ORed into
/// `acc` at `unit` (predicated ORs, no integer selects)
__device__ __forceinline__ void alpha_count_bits(uint32_t &acc, float a, float T0, float T1, float T2, float T3, float T4, float T5, float T6,
                                                 uint32_t unit)
{
        asm("{\n\t"
            ".reg .pred p1, p2, p3;\n\t"
            ".reg .f32 t, u;\n\t"
            "setp.le.f32 p1, %1, %5;\n\t"   // a <= T3
            "selp.f32 t, %7, %3, p1;\n\t"   // T5 : T1
            "setp.le.f32 p2, %1, t;\n\t"
            "selp.f32 t, %8, %6, p2;\n\t"   // p1: T6 : T4
            "selp.f32 u, %4, %2, p2;\n\t"   // !p1: T2 : T0
            "selp.f32 t, t, u, p1;\n\t"
            "setp.le.f32 p3, %1, t;\n\t"
            "@p1 or.b32 %0, %0, %9;\n\t"
            "@p2 or.b32 %0, %0, %10;\n\t"
            "@p3 or.b32 %0, %0, %11;\n\t"
            "}"
            : "+r"(acc)
            : "f"(a), "f"(T0), "f"(T1), "f"(T2), "f"(T3), "f"(T4), "f"(T5), "f"(T6), "r"(unit * 4u), "r"(unit * 2u), "r"(unit));
}

/// index = 1 + count, & 7, ^ (2 > index) (cuda_dxt.cu:376-388), i.e. 0 -> 0, 1..6 -> 2..7, 7 -> 1, on all 3-bit fields of a word at once
__device__ __forceinline__ uint32_t alpha_count_to_index(uint32_t x)
{
        constexpr uint32_t kLsb = 0x09249249u;  // bit 0 of each of the ten fields
        const uint32_t x1 = x >> 1, x2 = x >> 2;
        const uint32_t all7 = x & x1 & x2 & kLsb;           // fields equal to 7
        const uint32_t inc = (x | x1 | x2) & kLsb & ~all7;  // non-zero fields below 7 get + 1 (6 + 1 does not carry out of the field)
        return ((x & ~(all7 * 7u)) + inc) | all7;
}

/// dxt_encode<6>, cuda_dxt.cu:471-509 with helpers :141-410
__device__ __forceinline__ uint4 dxt6_encode(const float (&r)[16], const float (&g)[16], const float (&b)[16])
{
        const double offd = (double) kOffset;
        float Y[16];
        float2 Co[8], Cg[8];  // pixels 2 j (.x) and 2 j + 1 (.y): the pairing of the packed instructions below
        // ConvertRGBToYCoCg (:141-148): unsuffixed literals make these double expressions, narrowed once.  As compiled:
        //   Y  = ((r + 2 g) + b) * 0.25,  Co = fma((2 r - 2 b), 0.25, off),  Cg = fma(((-r + 2 g) - b), 0.25, off)
        // with g2 = g + g.  Written here with fewer FP64 instructions, each step the same real number rounded once: r + g2 = fma(g, 2, r)
        // (2 g is exact), 2 r - 2 b = 2 (r - b) exactly (scaling by 2 commutes with rounding) and fma(2 d, 0.25, off) = fma(d, 0.5, off).
#pragma unroll
        for (int i = 0; i < 16; ++i) {
                const double dr = (double) r[i], dg = (double) g[i], db = (double) b[i];
                Y[i] = __double2float_rn(__dmul_rn(__dadd_rn(__fma_rn(dg, 2.0, dr), db), 0.25));
                const float co = __double2float_rn(__fma_rn(__dadd_rn(dr, -db), 0.5, offd));
                const float cg = __double2float_rn(__fma_rn(__dadd_rn(__fma_rn(dg, 2.0, -dr), -db), 0.25, offd));
                if (i & 1) {
                        Co[i >> 1].y = co, Cg[i >> 1].y = cg;
                } else {
                        Co[i >> 1].x = co, Cg[i >> 1].x = cg;
                }
        }
        // FindMinMaxColorsBox (:159-168)
        float mnY = Y[0], mxY = Y[0], mnCo = Co[0].x, mxCo = Co[0].x, mnCg = Cg[0].x, mxCg = Cg[0].x;
#pragma unroll
        for (int i = 1; i < 16; ++i) {
                mnY = fminf(mnY, Y[i]), mxY = fmaxf(mxY, Y[i]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
                mnCo = fminf(mnCo, fminf(Co[j].x, Co[j].y)), mxCo = fmaxf(mxCo, fmaxf(Co[j].x, Co[j].y));
                mnCg = fminf(mnCg, fminf(Cg[j].x, Cg[j].y)), mxCg = fmaxf(mxCg, fmaxf(Cg[j].x, Cg[j].y));
        }
        // SelectYCoCgDiagonal (:260-270): t = c - (max+min)*0.5 is fma(max+min, -0.5, c); cov sequential from +0
        {
                const float sCo = __fadd_rn(mnCo, mxCo), sCg = __fadd_rn(mnCg, mxCg);
                float cov = 0.0f;
                const float2 so2 = dup(sCo), sg2 = dup(sCg), mh = dup(-0.5f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // the two deviations of pixels 2 j, 2 j + 1 packed; the chain itself stays sequential
                        const float2 eo = __ffma2_rn(so2, mh, Co[j]), eg = __ffma2_rn(sg2, mh, Cg[j]);
                        cov = __fmaf_rn(eo.x, eg.x, cov);
                        cov = __fmaf_rn(eo.y, eg.y, cov);
                }
                if (cov < 0.0f) {  // :485-489
                        const float t = mxCg;
                        mxCg = mnCg, mnCg = t;
                }
        }
        // ScaleYCoCg (:241-258)
        const float eXo = __fadd_rn(mxCo, -kOffset), eXg = __fadd_rn(mxCg, -kOffset);
        const float eNo = __fadd_rn(mnCo, -kOffset), eNg = __fadd_rn(mnCg, -kOffset);
        const float m = fmaxf(fmaxf(fabsf(eNo), fabsf(eNg)), fmaxf(fabsf(eXo), fabsf(eXg)));
        uint32_t scale = 1u;
        if (m < 0.2509804069995880127f) {  // (float)(64.0/255.0)
                scale = 2u;
        }
        if (m < 0.12549020349979400635f) {  // (float)(32.0/255.0)
                scale = 4u;
        }
        const float fs = (float) scale, inv_s = scale == 1u ? 1.0f : scale == 2u ? 0.5f : 0.25f;  // rcp.rn of 1,2,4 is exact

        // EmitEndPointsYCoCgDXT5 (:272-313)
        const float sXo = __fmaf_rn(eXo, fs, kOffset), sXg = __fmaf_rn(eXg, fs, kOffset);  // (c - off)*scale + off
        const float sNo = __fmaf_rn(eNo, fs, kOffset), sNg = __fmaf_rn(eNg, fs, kOffset);
        // InsetCoCgBBox (:182-187): (max-min)*(1/16) - (float)((8/255)/16) in one FMA
        const float kIns = (float) ((8.0 / 255.0) / 16.0);  // 0.0019607844...
        const float insO = __fmaf_rn(__fadd_rn(sXo, -sNo), 0.0625f, -kIns), insG = __fmaf_rn(__fadd_rn(sXg, -sNg), 0.0625f, -kIns);
        const float cXo = add_sat_rn(sXo, -insO), cXg = add_sat_rn(sXg, -insG);  // clamp(max - inset, 0, 1)
        const float cNo = add_sat_rn(sNo, insO), cNg = add_sat_rn(sNg, insG);    // clamp(min + inset, 0, 1)
        const uint32_t qXo = magic_bits(roundu_magic(__fmul_rn(cXo, 31.0f))), qXg = magic_bits(roundu_magic(__fmul_rn(cXg, 63.0f)));
        const uint32_t qNo = magic_bits(roundu_magic(__fmul_rn(cNo, 31.0f))), qNg = magic_bits(roundu_magic(__fmul_rn(cNg, 63.0f)));
        uint4 outp;
        outp.z = ((qXo << 11) | (qXg << 5) | (scale - 1u)) | (((qNo << 11) | (qNg << 5) | (scale - 1u)) << 16);
        // expand to 8 bits, back to unit range, undo the scale: fma(fma(float(e), 1/255, -off), 1/scale, off)
        const float k255 = 0.0039215688593685626984f;  // (float)(1.0/255.0)
#define UGB_EXPAND(q5or6, e)                                                                                           \
        __fmaf_rn(__fmaf_rn((float) (e), k255, -kOffset), inv_s, kOffset)
        const float pXo = UGB_EXPAND(qXo, (qXo << 3) | (qXo >> 2)), pXg = UGB_EXPAND(qXg, (qXg << 2) | (qXg >> 4));
        const float pNo = UGB_EXPAND(qNo, (qNo << 3) | (qNo >> 2)), pNg = UGB_EXPAND(qNg, (qNg << 2) | (qNg >> 4));
#undef UGB_EXPAND

        // EmitIndicesYCoCgDXT5 (:315-348).  Palette c0 = max, c1 = min, c2/c3 = lerp with (float)(1/3), (float)(2/3);
        // which product is the plain multiply and which rides the FMA differs between Co and Cg (as compiled).
        const float c2o = __fmaf_rn(pNo, 0.3333333432674407959f, __fmul_rn(pXo, 0.66666662693023681641f));
        const float c2g = __fmaf_rn(pXg, 0.66666662693023681641f, __fmul_rn(pNg, 0.3333333432674407959f));
        const float c3o = __fmaf_rn(pXo, 0.3333333134651184082f, __fmul_rn(pNo, 0.6666666865348815918f));
        const float c3g = __fmaf_rn(pXg, 0.3333333134651184082f, __fmul_rn(pNg, 0.6666666865348815918f));
        // colorDistance = fma(dCo, dCo, dCg*dCg) for the four palette entries; pixels i and i+1 share packed (f32x2) instructions —
        // same operations, half the issue slots
        uint32_t cidx = 0;
        const float2 nXo = dup(-pXo), nXg = dup(-pXg), nNo = dup(-pNo), nNg = dup(-pNg), n2o = dup(-c2o), n2g = dup(-c2g), n3o = dup(-c3o),
                     n3g = dup(-c3g);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
                const float2 co = Co[i >> 1], cg = Cg[i >> 1];
#define UGB_DIST2(no, ng, d)                                                                                                               \
        {                                                                                                                                  \
                const float2 eo = __fadd2_rn(co, no), eg = __fadd2_rn(cg, ng);                                                             \
                d = __ffma2_rn(eo, eo, __fmul2_rn(eg, eg));                                                                                \
        }
                float2 d0, d1, d2, d3;
                UGB_DIST2(nXo, nXg, d0)
                UGB_DIST2(nNo, nNg, d1)
                UGB_DIST2(n2o, n2g, d2)
                UGB_DIST2(n3o, n3g, d3)
#undef UGB_DIST2
                color_index_bits(cidx, d0.x, d1.x, d2.x, d3.x, 1u << (2 * i), 2u << (2 * i));
                color_index_bits(cidx, d0.y, d1.y, d2.y, d3.y, 1u << (2 * i + 2), 2u << (2 * i + 2));
        }
        outp.w = cidx;

        // InsetYBBox (:176-181): (max - min)/32.0 - (16.0/255.0)/32.0 in double, narrowed once
        const float insY = __double2float_rn(__fma_rn((double) __fadd_rn(mxY, -mnY), 1.0 / 32.0, -((16.0 / 255.0) / 32.0)));
        const float nY = add_sat_rn(mnY, insY), xY = add_sat_rn(mxY, -insY);
        // EmitAlphaEndPointsYCoCgDXT5 (:350-357): roundf(c * 255.0); the double product narrows to the float product
        const uint32_t a0 = magic_bits(roundu_magic(__fmul_rn(nY, 255.0f))), a1 = magic_bits(roundu_magic(__fmul_rn(xY, 255.0f)));
        // EmitAlphaIndicesYCoCgDXT5 (:360-410)
        const float mid = __fdiv_rn(__fadd_rn(xY, -nY), 14.0f);  // (max-min)/(2.0*7.0): float division is what was compiled
        const double dX = (double) xY, dN = (double) nY, dM = (double) mid;
        const double k7 = 1.0 / 7.0;
        float ab[7];
        ab[0] = __fadd_rn(nY, mid);
        ab[1] = __double2float_rn(__fma_rn(__fma_rn(dX, 6.0, dN), k7, dM));
        ab[2] = __double2float_rn(__fma_rn(__fma_rn(dX, 5.0, __dadd_rn(dN, dN)), k7, dM));
        ab[3] = __double2float_rn(__fma_rn(__fma_rn(dX, 4.0, __dmul_rn(dN, 3.0)), k7, dM));
        ab[4] = __double2float_rn(__fma_rn(__fma_rn(dX, 3.0, __dmul_rn(dN, 4.0)), k7, dM));
        ab[5] = __double2float_rn(__fma_rn(__fma_rn(dX, 2.0, __dmul_rn(dN, 5.0)), k7, dM));
        ab[6] = __double2float_rn(__fma_rn(__fma_rn(dN, 6.0, dX), k7, dM));
        // index = 1 + #{k : Y <= ab_k}, & 7, ^ (2 > index)  (:376-388).  The thresholds are ordered ab2 >= ab3 >= ... >= ab7 >= ab1
        // (rounding is monotone and max >= min), so the count is a 3-step binary search instead of 7 compares, and the
        // "& 7, ^ (2 > idx)" fix-up is the map 0,2,3,4,5,6,7,1 of the count (alpha_count_to_index).
        const float T0 = ab[1], T1 = ab[2], T2 = ab[3], T3 = ab[4], T4 = ab[5], T5 = ab[6], T6 = ab[0];
        // count fields (3 bits per pixel) of pixels 0..9 in cntA, 10..15 in cntB; the map count -> index runs once per word afterwards
        uint32_t cntA = 0, cntB = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
                alpha_count_bits(i < 10 ? cntA : cntB, Y[i], T0, T1, T2, T3, T4, T5, T6, 1u << (3 * (i < 10 ? i : i - 10)));
        }
        const uint32_t idxA = alpha_count_to_index(cntA), idxB = alpha_count_to_index(cntB);
        // the 48-bit index string (pixel i at bit 3 i) follows the two endpoint bytes (:389-392)
        const uint32_t s_lo = idxA | (idxB << 30), s_hi = idxB >> 2;
        outp.x = (a0 << 8) | a1 | (s_lo << 16);
        outp.y = (s_lo >> 16) | (s_hi << 16);
        return outp;
}

}  // namespace ugb
