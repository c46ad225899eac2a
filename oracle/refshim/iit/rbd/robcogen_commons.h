// refshim/iit/rbd/robcogen_commons.h — see rbd.h.  The spatial-algebra helpers the generated ForwardDynamics::fd calls
// (/root/reference/src/auvsl_dynamics/generated/forward_dynamics.cpp:62-66, 105, 116-120, ...).
#pragma once
#include "rbd.h"
namespace iit {
namespace rbd {

// [[ w^, 0 ], [ l^, w^ ]] : matrix of the spatial cross product for motion vectors, v = (w; l)
template <typename Scalar> inline void motionCrossProductMx(const Vec6<Scalar>& v, Mat66<Scalar>& mx) {
    mx.setZero();
    mx(AX, AY) = -v(AZ); mx(AX, AZ) =  v(AY);
    mx(AY, AX) =  v(AZ); mx(AY, AZ) = -v(AX);
    mx(AZ, AX) = -v(AY); mx(AZ, AY) =  v(AX);
    mx(LX, AY) = -v(LZ); mx(LX, AZ) =  v(LY);
    mx(LY, AX) =  v(LZ); mx(LY, AZ) = -v(LX);
    mx(LZ, AX) = -v(LY); mx(LZ, AY) =  v(LX);
    mx(LX, LY) = -v(AZ); mx(LX, LZ) =  v(AY);
    mx(LY, LX) =  v(AZ); mx(LY, LZ) = -v(AX);
    mx(LZ, LX) = -v(AY); mx(LZ, LY) =  v(AX);
}
template <typename Scalar> inline Mat66<Scalar> motionCrossProductMx(const Vec6<Scalar>& v) { Mat66<Scalar> m; motionCrossProductMx<Scalar>(v, m); return m; }
template <typename Scalar> inline void forceCrossProductMx(const Vec6<Scalar>& v, Mat66<Scalar>& mx) {
    Mat66<Scalar> m; motionCrossProductMx<Scalar>(v, m);
    mx = -(m.transpose());
}

// v x* (I v)
template <typename Scalar, int R2, int C2> inline Vec6<Scalar> vxIv(const Vec6<Scalar>& v, const Eigen::Matrix<Scalar, R2, C2>& I) {
    Vec6<Scalar> Iv = Mat66<Scalar>(I) * v;
    Vec6<Scalar> r;
    r(AX) = v(AY) * Iv(AZ) - v(AZ) * Iv(AY) + v(LY) * Iv(LZ) - v(LZ) * Iv(LY);
    r(AY) = v(AZ) * Iv(AX) - v(AX) * Iv(AZ) + v(LZ) * Iv(LX) - v(LX) * Iv(LZ);
    r(AZ) = v(AX) * Iv(AY) - v(AY) * Iv(AX) + v(LX) * Iv(LY) - v(LY) * Iv(LX);
    r(LX) = v(AY) * Iv(LZ) - v(AZ) * Iv(LY);
    r(LY) = v(AZ) * Iv(LX) - v(AX) * Iv(LZ);
    r(LZ) = v(AX) * Iv(LY) - v(AY) * Iv(LX);
    return r;
}

// articulated inertia of a link attached through a revolute joint (axis AZ): Ia = IA - U U^T / D; row / column AZ of the
// result are analytically zero and are not written (the caller zero-initialises Ia once)
template <typename Scalar, int R2, int C2> inline void compute_Ia_revolute(const Eigen::Matrix<Scalar, R2, C2>& IA, const Vec6<Scalar>& U, const Scalar& D, Mat66<Scalar>& Ia) {
    for (int i = 0; i < 6; i++) {
        if (i == AZ) continue;
        const Scalar UD = U(i) / D;
        for (int j = 0; j < 6; j++) {
            if (j == AZ) continue;
            Ia(i, j) = IA(i, j) - UD * U(j);
        }
    }
}
// the same for a prismatic joint (axis LZ)
template <typename Scalar, int R2, int C2> inline void compute_Ia_prismatic(const Eigen::Matrix<Scalar, R2, C2>& IA, const Vec6<Scalar>& U, const Scalar& D, Mat66<Scalar>& Ia) {
    for (int i = 0; i < 6; i++) {
        if (i == LZ) continue;
        const Scalar UD = U(i) / D;
        for (int j = 0; j < 6; j++) {
            if (j == LZ) continue;
            Ia(i, j) = IA(i, j) - UD * U(j);
        }
    }
}
// coordinate transform of an articulated inertia to the parent frame: X^T Ia X  (X: motion transform child <- parent)
template <typename Scalar, int R2, int C2> inline void ctransform_Ia_revolute(const Mat66<Scalar>& Ia, const Eigen::Matrix<Scalar, R2, C2>& XM, Mat66<Scalar>& Ia_B) {
    const Mat66<Scalar> X(XM);
    Ia_B = X.transpose() * (Ia * X);
}
template <typename Scalar, int R2, int C2> inline void ctransform_Ia_prismatic(const Mat66<Scalar>& Ia, const Eigen::Matrix<Scalar, R2, C2>& XM, Mat66<Scalar>& Ia_B) {
    ctransform_Ia_revolute<Scalar>(Ia, XM, Ia_B);
}

}  // namespace rbd
}  // namespace iit
