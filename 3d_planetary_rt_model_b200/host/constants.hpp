// constants.hpp -- physical constants of the host facade; values as the reference's src/constants.hpp:9-63
#pragma once
#include <cmath>
#include <string>

namespace b200rt_host {

typedef double Real;   // the facade computes its inputs in double, like the reference CPU build (Real.hpp:18-27)

constexpr Real rMars = 3395e5;                 // cm
constexpr Real mMars = 0.1076 * 5.98e27;       // g
constexpr Real G = 6.67e-8;                    // dyn cm^2 g^-2
constexpr Real kB = 1.38e-16;                  // erg K^-1
constexpr Real clight = 3e10;                  // cm s^-1
constexpr Real mH = 1.673e-24;                 // g
constexpr Real mCO2 = 44 * mH;
constexpr Real line_f_coeff = 2.647e-2;        // cm^2 Hz
constexpr Real aMars_typical = 1.41;           // AU
constexpr Real pi = M_PI;
constexpr Real rexo_typical = rMars + 200e5;   // typical exobase altitude (constants.hpp:12)

constexpr Real lyman_alpha_lambda = 121.6e-7;  // cm
constexpr Real lyman_alpha_f = 0.41641;
const Real lyman_alpha_cross_section_total = line_f_coeff * lyman_alpha_f;
const Real lyman_alpha_line_center_cross_section_coef =
    lyman_alpha_cross_section_total / std::sqrt(2.0 * pi * kB / mH) * lyman_alpha_lambda;
constexpr Real CO2_lyman_alpha_absorption_cross_section = 6.3e-20;
const Real lyman_alpha_flux_Earth_typical = 4.5e15;
const Real lyman_alpha_flux_Mars_typical = (lyman_alpha_flux_Earth_typical / 1e4 * 1e8 * lyman_alpha_lambda *
                                            lyman_alpha_lambda / clight / aMars_typical / aMars_typical);
const Real lyman_alpha_typical_g_factor = lyman_alpha_flux_Mars_typical * lyman_alpha_cross_section_total;

constexpr Real lyman_beta_lambda = 102.6e-7;
constexpr Real lyman_beta_f = 0.079142;
const Real lyman_beta_cross_section_total = line_f_coeff * lyman_beta_f;
const Real lyman_beta_branching_ratio = 0.8819;
const Real lyman_beta_line_center_cross_section_coef =
    lyman_beta_cross_section_total / std::sqrt(2.0 * pi * kB / mH) * lyman_beta_lambda;
constexpr Real CO2_lyman_beta_absorption_cross_section = 3.53e-17;
const Real lyman_beta_flux_Earth_typical = lyman_alpha_flux_Earth_typical / 66.;
const Real lyman_beta_flux_Mars_typical = (lyman_beta_flux_Earth_typical / 1e4 * 1e8 * lyman_beta_lambda *
                                           lyman_beta_lambda / clight / aMars_typical / aMars_typical);
const Real lyman_beta_typical_g_factor = lyman_beta_flux_Mars_typical * lyman_beta_cross_section_total;

} // namespace b200rt_host
