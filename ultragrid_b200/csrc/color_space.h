// Q14 integer RGB<->YCbCr coefficients — the contract of UltraGrid's src/color_space.{h,c}.
//
// The reference derives the coefficients in double at compile time (color_space.c:46-131) and rounds
// with +-0.5 before truncating to int (C_EPS, color_space.c:55).  The same formulas are evaluated here
// as constexpr and pinned with static_asserts to the values probed from the reference build
// (SURVEY.md section 8 A3), so a transcription error cannot compile.
#pragma once
#include <stdint.h>

namespace ugb {

constexpr int COMP_BASE = 14;  // color_space.h:70 (comp_type_t is int32_t)

constexpr double KR_709 = .212639, KB_709 = .072192;  // color_space.h:75-76
constexpr double KR_601 = .299, KB_601 = .114;        // color_space.h:73-74

struct color_coeffs {  // field meaning as color_space.h:135-149 (all widened to int here)
        int y_r, y_g, y_b;
        int cb_r, cb_g, cb_b;
        int cr_r, cr_g, cr_b;
        int y_scale;
        int r_cr, g_cb, g_cr;
        int b_cb;
};

namespace detail {
constexpr double kg(double kr, double kb) { return 1. - kr - kb; }
constexpr double dd(double kr, double kb) { return 2. * (kr + kg(kr, kb)); }  // D(), color_space.c:47
constexpr double ee(double kr) { return 2. * (1. - kr); }                     // E(), color_space.c:48
// limited-range scale factors, color_space.c:56-63; depth 0 = full range
constexpr double y_limit(int d) { return d == 0 ? 1.0 : (219. * (1 << (d - 8)) / ((1 << d) - 1)); }
constexpr double c_limit(int d) { return d == 0 ? 1.0 : (224. * (1 << (d - 8)) / ((1 << d) - 1)); }
constexpr int    to_i(double v) { return (int) v; }  // C cast: truncation toward zero
constexpr int    scaled(double x) { return to_i(x * (1 << COMP_BASE) + (x > 0 ? 1. : -1.) * 0.5); }  // color_space.c:104-105
}  // namespace detail

/// compute_color_coeffs(), color_space.c:192-196 / COEFFS(), color_space.c:116-128
constexpr color_coeffs compute_color_coeffs(double kr, double kb, int depth)
{
        using namespace detail;
        const double B = 1 << COMP_BASE;
        return color_coeffs{
                to_i(kr * y_limit(depth) * B + 0.5),
                to_i(kg(kr, kb) * y_limit(depth) * B + 0.5),
                to_i(kb * y_limit(depth) * B + 0.5),
                to_i(-kr / dd(kr, kb) * c_limit(depth) * B - 0.5),
                to_i(-kg(kr, kb) / dd(kr, kb) * c_limit(depth) * B - 0.5),
                to_i((1 - kb) / dd(kr, kb) * c_limit(depth) * B + 0.5),
                to_i((1 - kr) / ee(kr) * c_limit(depth) * B - 0.5),
                to_i(-kg(kr, kb) / ee(kr) * c_limit(depth) * B - 0.5),
                to_i(-kb / ee(kr) * c_limit(depth) * B + 0.5),
                scaled(1. / y_limit(depth)),
                scaled((2. * (1. - kr)) / c_limit(depth)),
                scaled((-kb * (2. * (kr + kg(kr, kb))) / kg(kr, kb)) / c_limit(depth)),
                scaled((-kr * (2. * (1. - kr)) / kg(kr, kb)) / c_limit(depth)),
                scaled((2. * (kr + kg(kr, kb))) / c_limit(depth)),
        };
}

/// get_color_coeffs(CS_DFL, depth) with the default (BT.709) colour space, color_space.c:149-183.
/// depth 0 = full range.
constexpr color_coeffs coeffs_709(int depth) { return compute_color_coeffs(KR_709, KB_709, depth); }
constexpr color_coeffs coeffs_601(int depth) { return compute_color_coeffs(KR_601, KB_601, depth); }

// ---- pinned to the reference build --------------------------------------------
namespace pin {
constexpr color_coeffs c8 = coeffs_709(8), c10 = coeffs_709(10), c16 = coeffs_709(16), c0 = coeffs_709(0);
static_assert(c8.y_r == 2992 && c8.y_g == 10063 && c8.y_b == 1016, "709/8 Y row");
static_assert(c8.cb_r == -1649 && c8.cb_g == -5547 && c8.cb_b == 7196, "709/8 Cb row");
static_assert(c8.cr_r == 7195 && c8.cr_g == -6536 && c8.cr_b == -659, "709/8 Cr row");
static_assert(c8.y_scale == 19077 && c8.r_cr == 29371 && c8.g_cb == -3494 && c8.g_cr == -8733 && c8.b_cb == 34610, "709/8 inverse");
static_assert(c10.y_r == 2983 && c10.y_g == 10034 && c10.y_b == 1013, "709/10 Y row");
static_assert(c10.cb_r == -1644 && c10.cb_g == -5531 && c10.cb_b == 7175, "709/10 Cb row");
static_assert(c10.cr_r == 7174 && c10.cr_g == -6517 && c10.cr_b == -657, "709/10 Cr row");
static_assert(c10.y_scale == 19133 && c10.r_cr == 29457 && c10.g_cb == -3504 && c10.g_cr == -8758 && c10.b_cb == 34712, "709/10 inverse");
static_assert(c16.y_r == 2980 && c16.y_g == 10024 && c16.y_b == 1012, "709/16 Y row");
static_assert(c16.cb_r == -1643 && c16.cb_g == -5525 && c16.cb_b == 7168, "709/16 Cb row");
static_assert(c16.cr_r == 7167 && c16.cr_g == -6511 && c16.cr_b == -656, "709/16 Cr row");
static_assert(c16.y_scale == 19152 && c16.r_cr == 29486 && c16.g_cb == -3507 && c16.g_cr == -8767 && c16.b_cb == 34745, "709/16 inverse");
static_assert(c0.y_r == 3484 && c0.y_g == 11717 && c0.y_b == 1183, "709/full Y row");
static_assert(c0.cb_r == -1877 && c0.cb_g == -6315 && c0.cb_b == 8192, "709/full Cb row");
static_assert(c0.cr_r == 8191 && c0.cr_g == -7441 && c0.cr_b == -750, "709/full Cr row");
static_assert(c0.y_scale == 16384 && c0.r_cr == 25800 && c0.g_cb == -3069 && c0.g_cr == -7671 && c0.b_cb == 30402, "709/full inverse");
}  // namespace pin

}  // namespace ugb
