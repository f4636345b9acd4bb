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

/// dxt_encode<6>, cuda_dxt.cu:471-509 with helpers :141-410
__device__ __forceinline__ uint4 dxt6_encode(const float (&r)[16], const float (&g)[16], const float (&b)[16])
{
        const double offd = (double) kOffset;
        float Y[16], Co[16], Cg[16];
        // ConvertRGBToYCoCg (:141-148): unsuffixed literals make these double expressions, narrowed once
#pragma unroll
        for (int i = 0; i < 16; ++i) {
                const double dr = (double) r[i], dg = (double) g[i], db = (double) b[i];
                const double g2 = __dadd_rn(dg, dg);
                Y[i] = __double2float_rn(__dmul_rn(__dadd_rn(__dadd_rn(dr, g2), db), 0.25));
                Co[i] = __double2float_rn(__fma_rn(__dadd_rn(__dadd_rn(dr, dr), -__dadd_rn(db, db)), 0.25, offd));
                Cg[i] = __double2float_rn(__fma_rn(__dadd_rn(__dadd_rn(-dr, g2), -db), 0.25, offd));
        }
        // FindMinMaxColorsBox (:159-168)
        float mnY = Y[0], mxY = Y[0], mnCo = Co[0], mxCo = Co[0], mnCg = Cg[0], mxCg = Cg[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) {
                mnY = fminf(mnY, Y[i]), mxY = fmaxf(mxY, Y[i]);
                mnCo = fminf(mnCo, Co[i]), mxCo = fmaxf(mxCo, Co[i]);
                mnCg = fminf(mnCg, Cg[i]), mxCg = fmaxf(mxCg, Cg[i]);
        }
        // SelectYCoCgDiagonal (:260-270): t = c - (max+min)*0.5 is fma(max+min, -0.5, c); cov sequential from +0
        {
                const float sCo = __fadd_rn(mnCo, mxCo), sCg = __fadd_rn(mnCg, mxCg);
                float cov = 0.0f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                        cov = __fmaf_rn(__fmaf_rn(sCo, -0.5f, Co[i]), __fmaf_rn(sCg, -0.5f, Cg[i]), cov);
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
                const float2 co = f2(Co[i], Co[i + 1]), cg = f2(Cg[i], Cg[i + 1]);
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
                {
                        const uint32_t bx = d0.x > d3.x, by = d1.x > d2.x, bz = d0.x > d2.x, bw = d1.x > d3.x, b4 = d2.x > d3.x;
                        cidx |= ((bx & b4) | (((by & bz) | (bx & bw)) << 1)) << (2 * i);
                }
                {
                        const uint32_t bx = d0.y > d3.y, by = d1.y > d2.y, bz = d0.y > d2.y, bw = d1.y > d3.y, b4 = d2.y > d3.y;
                        cidx |= ((bx & b4) | (((by & bz) | (bx & bw)) << 1)) << (2 * i + 2);
                }
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
        // "& 7, ^ (2 > idx)" fix-up is the nibble table 0,2,3,4,5,6,7,1 indexed by the count.
        const float T0 = ab[1], T1 = ab[2], T2 = ab[3], T3 = ab[4], T4 = ab[5], T5 = ab[6], T6 = ab[0];
        uint32_t ix = 0, iy = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
                const float a = Y[i];
                const bool p1 = a <= T3;
                const bool p2 = a <= (p1 ? T5 : T1);
                const bool p3 = a <= (p1 ? (p2 ? T6 : T4) : (p2 ? T2 : T0));
                const uint32_t cnt = (p1 ? 4u : 0u) + (p2 ? 2u : 0u) + (p3 ? 1u : 0u);
                const uint32_t idx = (0x17654320u >> (4u * cnt)) & 7u;
                if (i < 6) {
                        ix |= idx << (3 * i + 16);  // pixel 5 keeps only its bit 0 here (:389) ...
                }
                if (i == 5) {
                        iy = idx >> 1;  // ... and the rest opens the second word (:392)
                }
                if (i > 5) {
                        iy |= idx << (3 * i - 16);
                }
        }
        outp.x = (a0 << 8) | a1 | ix;
        outp.y = iy;
        return outp;
}

}  // namespace ugb
