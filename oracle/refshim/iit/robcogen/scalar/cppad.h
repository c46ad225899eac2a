// refshim/iit/robcogen/scalar/cppad.h — see iit/rbd/rbd.h.  Scalar traits for CppAD::AD<double> (generated/rbd_types.h:14).
#pragma once
#include <cppad/cppad.hpp>
namespace iit {
namespace robcogen {
struct CppADDoubleTraits {
    typedef CppAD::AD<double> Scalar;
    typedef double ValueType;
    static Scalar sin(const Scalar& x) { return CppAD::sin(x); }
    static Scalar cos(const Scalar& x) { return CppAD::cos(x); }
    static Scalar tan(const Scalar& x) { return CppAD::tan(x); }
    static Scalar sqrt(const Scalar& x) { return CppAD::sqrt(x); }
    static Scalar abs(const Scalar& x) { return CppAD::abs(x); }
};
}  // namespace robcogen
}  // namespace iit
