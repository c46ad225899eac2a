// refshim/iit/rbd/InertiaMatrix.h — see rbd.h.
#pragma once
#include "rbd.h"
namespace iit {
namespace rbd {
// 6x6 spatial inertia about the link-frame origin: [[ I, m c^ ], [ -m c^, m 1 ]]
template <typename SCALAR> class InertiaMat : public Core<SCALAR>::Matrix66 {
public:
    typedef typename Core<SCALAR>::Matrix66 Base;
    typedef SCALAR Scalar;
    typedef typename Core<SCALAR>::Vector3 Vec3;
    typedef typename Core<SCALAR>::Matrix33 Mat33;
    InertiaMat() { this->setZero(); }
    InertiaMat(const Scalar& m, const Vec3& com, const Mat33& tensor) { this->setZero(); fill(m, com, tensor); }
    void fill(const Scalar& mass, const Vec3& c, const Mat33& tensor) {
        Base& M = *this;
        M.setZero();
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M(AX + i, AX + j) = tensor(i, j);
        const Scalar mx = mass * c[0], my = mass * c[1], mz = mass * c[2];
        Mat33 mc;  // m c^
        mc(0, 0) = Scalar(0); mc(0, 1) = -mz;       mc(0, 2) = my;
        mc(1, 0) = mz;        mc(1, 1) = Scalar(0); mc(1, 2) = -mx;
        mc(2, 0) = -my;       mc(2, 1) = mx;        mc(2, 2) = Scalar(0);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { M(AX + i, LX + j) = mc(i, j); M(LX + i, AX + j) = -mc(i, j); }
        for (int i = 0; i < 3; i++) M(LX + i, LX + i) = mass;
    }
    Scalar getMass() const { return (*this)(LX, LX); }
    Vec3 getCOM() const {
        const Scalar m = getMass();
        return Vec3((*this)(AZ, LY) / m, (*this)(AX, LZ) / m, (*this)(AY, LX) / m);
    }
    Mat33 get3x3Tensor() const { return this->template block<3, 3>(AX, AX); }
    template <int R2, int C2> InertiaMat& operator=(const Eigen::Matrix<SCALAR, R2, C2>& o) { Base::operator=(o); return *this; }
};
}  // namespace rbd
}  // namespace iit
