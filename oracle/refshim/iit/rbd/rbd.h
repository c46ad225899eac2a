// refshim/iit/rbd/rbd.h — TEST INFRASTRUCTURE ONLY (oracle/README.md).
//
// Stand-in for the iit-rbd headers (the run-time library of RobCoGen-generated code: https://github.com/mfrigerio17/iit-rbd)
// that the reference's generated sources include.  iit-rbd is an un-vendored, un-pinned dependency of the reference
// (README.md:8) and is absent from this image; what follows restates the PUBLISHED definitions its generated code relies on:
//   * spatial vectors are [angular(3); linear(3)], coordinates AX, AY, AZ, LX, LY, LZ = 0..5
//   * InertiaMat::fill(m, c, I) = [[ I, m c^ ], [ -m c^, m 1 ]]  with  c^ = [[0,-cz,cy],[cz,0,-cx],[-cy,cx,0]]
//   * Utils::buildInertiaTensor(Ixx,Iyy,Izz,Ixy,Ixz,Iyz) = [[Ixx,-Ixy,-Ixz],[-Ixy,Iyy,-Iyz],[-Ixz,-Iyz,Izz]]
//   * Utils::transform(H, p) = H(0:3,0:3) p + H(0:3,3)
//   * motionCrossProductMx(v) = [[w^, 0],[l^, w^]],  vxIv(v, I) = v x* (I v) = [ w x n + l x f ; w x f ],  (n; f) = I v
//   * compute_Ia_revolute(IA, U, D, Ia): Ia = IA - U U^T / D, row and column AZ left untouched (they are analytically 0)
//   * ctransform_Ia_revolute(Ia, X, IaB): IaB = X^T Ia X
// The real library exploits sparsity in some of these products; results can differ from it by rounding only.
#pragma once
#include <Eigen/Dense>

namespace iit {
namespace rbd {

template <typename SCALAR, int R, int C> using PlainMatrix = Eigen::Matrix<SCALAR, R, C>;

enum Coordinates6D { AX = 0, AY, AZ, LX, LY, LZ };
enum Coords3D { X = 0, Y, Z, W };

static const double g = 9.81;

template <typename SCALAR> struct Core {
    typedef SCALAR Scalar;
    typedef PlainMatrix<Scalar, 6, 6> Matrix66;
    typedef PlainMatrix<Scalar, 3, 3> Matrix33;
    typedef PlainMatrix<Scalar, 4, 4> Matrix44;
    typedef PlainMatrix<Scalar, 6, 1> Vector6;
    typedef PlainMatrix<Scalar, 6, 1> Column6D;
    typedef PlainMatrix<Scalar, 3, 1> Vector3;
    typedef Vector6 ForceVector;
    typedef Vector6 VelocityVector;
};

template <typename S> using Mat66 = PlainMatrix<S, 6, 6>;
template <typename S> using Mat33 = PlainMatrix<S, 3, 3>;
template <typename S> using Vec6 = PlainMatrix<S, 6, 1>;
template <typename S> using Vec3 = PlainMatrix<S, 3, 1>;

}  // namespace rbd
}  // namespace iit
