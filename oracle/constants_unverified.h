/*
 * oracle/constants_unverified.h  --  TEST INFRASTRUCTURE, not product code.
 *
 * Every constant in this header restates arithmetic that lives in third-party crates pinned by the
 * reference's Cargo.lock but NOT vendored under /root/reference (ssimulacra2 0.5.1, yuvxyb 0.4.2,
 * yuvxyb-math 0.1.1, palette 0.7.6, cogset 0.2.0).  They are written down from the published
 * algorithms (libjxl tools/ssimulacra2.cc + lib/jxl/gauss_blur.cc, zimg gamma.cpp, CIE 15:2004,
 * Sharma/Wu/Dalal 2005) and could not be diffed against the crates on this box:
 *
 *        >>>>>  PARITY UNPINNED against the real crates  <<<<<
 *
 * Keeping them in one header means a verified copy can replace them wholesale.  What IS pinned:
 * CIEDE2000 against the 34 published Sharma test pairs, Lab against D65 identities, the recursive
 * Gaussian against its closed-form FIR (sum of taps == 1), SSIMULACRA2(identical) == 100.
 */
#ifndef SNES_ORACLE_CONSTANTS_UNVERIFIED_H
#define SNES_ORACLE_CONSTANTS_UNVERIFIED_H

/* ---- yuvxyb: sRGB EOTF (H.273 / zimg constants) ------------------------------------------- */
#define ORA_SRGB_ALPHA 1.0550107f   /* 1.055010718947587 rounded to f32 */
#define ORA_SRGB_BETA  0.0030412825f /* 0.003041282560128 rounded to f32 */

/* ---- yuvxyb: linear RGB -> XYB (libjxl opsin absorbance) ------------------------------------ */
#define ORA_K_M02 0.078f
#define ORA_K_M00 0.30f
#define ORA_K_M01 (1.0f - ORA_K_M02 - ORA_K_M00)
#define ORA_K_M12 0.078f
#define ORA_K_M10 0.23f
#define ORA_K_M11 (1.0f - ORA_K_M12 - ORA_K_M10)
#define ORA_K_M20 0.24342269f
#define ORA_K_M21 0.20476745f
#define ORA_K_M22 (1.0f - ORA_K_M20 - ORA_K_M21)
#define ORA_K_B0 0.0037930734f
/* -cbrt(K_B0) rounded to f32 */
#define ORA_NEG_CBRT_BIAS (-0.15595420f)

/* ---- ssimulacra2: blur sigma, SSIM C2, pooling ---------------------------------------------- */
#define ORA_BLUR_SIGMA 1.5
#define ORA_SSIM_C2 0.0009f
#define ORA_NUM_SCALES 6

static const double ORA_SSIM2_WEIGHT[108] = {
    0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0, 0.0004371155730107379, 0.0,
    1.1041726426657346, 0.00066284834129271, 0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0,
    1.8422455520539298, 11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
    1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072, 0.9677937080626833, 0.0,
    0.00014003424285435884, 0.9981766977854967, 0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0,
    0.0013648766163243398, 0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296, 1.027889937768264,
    225.20515300849274, 0.0, 0.0, 19.213238186143016, 0.0011401524586618361, 0.001237755635509985,
    176.39317598450694, 0.0, 0.0, 24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
    34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0008680556573291698, 0.0, 0.0, 0.0,
    0.0, 0.0, 0.0005313191874358747, 0.0, 0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0004179171803251336,
    0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862, 23.19243343998926, 0.0,
    95.1080498811086, 0.9863978034400682, 0.9834382792465353, 0.0012286405048278493, 171.2667255897307,
    0.9807858872435379, 0.0, 0.0, 0.0, 0.0005130064588990679, 0.0, 0.00010854057858411537};

#define ORA_POOL_SCALE 0.9562382616834844
#define ORA_POOL_C1 2.326765642916932
#define ORA_POOL_C2 (-0.020884521182843837)
#define ORA_POOL_C3 6.248496625763138e-5
#define ORA_POOL_EXP 0.6276336467831387

/* ---- palette 0.7.6: sRGB <-> XYZ (D65) <-> Lab ---------------------------------------------- */
#define ORA_XYZ_M00 0.4124564
#define ORA_XYZ_M01 0.3575761
#define ORA_XYZ_M02 0.1804375
#define ORA_XYZ_M10 0.2126729
#define ORA_XYZ_M11 0.7151522
#define ORA_XYZ_M12 0.0721750
#define ORA_XYZ_M20 0.0193339
#define ORA_XYZ_M21 0.1191920
#define ORA_XYZ_M22 0.9503041
#define ORA_RGB_M00 3.2404542
#define ORA_RGB_M01 (-1.5371385)
#define ORA_RGB_M02 (-0.4985314)
#define ORA_RGB_M10 (-0.9692660)
#define ORA_RGB_M11 1.8760108
#define ORA_RGB_M12 0.0415560
#define ORA_RGB_M20 0.0556434
#define ORA_RGB_M21 (-0.2040259)
#define ORA_RGB_M22 1.0572252
#define ORA_D65_X 0.95047
#define ORA_D65_Y 1.0
#define ORA_D65_Z 1.08883

/* ---- cogset 0.2.0 k-means defaults ----------------------------------------------------------- */
#define ORA_KMEANS_TOL 1e-6
#define ORA_KMEANS_MAX_ITER 100

#endif
