// refshim/iit/rbd/utils.h — see rbd.h.
#pragma once
#include "rbd.h"
namespace iit {
namespace rbd {
class Utils {
public:
    // off-diagonal entries are the NEGATED products of inertia
    template <typename Scalar>
    static Mat33<Scalar> buildInertiaTensor(const Scalar& Ixx, const Scalar& Iyy, const Scalar& Izz, const Scalar& Ixy, const Scalar& Ixz, const Scalar& Iyz) {
        Mat33<Scalar> T;
        T << Ixx, -Ixy, -Ixz,
            -Ixy,  Iyy, -Iyz,
            -Ixz, -Iyz,  Izz;
        return T;
    }
    template <typename Scalar, typename A, typename B, typename C, typename D, typename E, typename F>
    static Mat33<Scalar> buildInertiaTensor(const A& Ixx, const B& Iyy, const C& Izz, const D& Ixy, const E& Ixz, const F& Iyz) {
        return buildInertiaTensor<Scalar>(Scalar(Ixx), Scalar(Iyy), Scalar(Izz), Scalar(Ixy), Scalar(Ixz), Scalar(Iyz));
    }
    // homogeneous transform applied to a point
    template <typename Scalar> static Vec3<Scalar> transform(const PlainMatrix<Scalar, 4, 4>& H, const Vec3<Scalar>& p) {
        Vec3<Scalar> r;
        for (int i = 0; i < 3; i++) r[i] = H(i, 0) * p[0] + H(i, 1) * p[1] + H(i, 2) * p[2] + H(i, 3);
        return r;
    }
    template <typename Scalar> static Vec3<Scalar> positionVector(const PlainMatrix<Scalar, 4, 4>& H) { return Vec3<Scalar>(H(0, 3), H(1, 3), H(2, 3)); }
};
}  // namespace rbd
}  // namespace iit
