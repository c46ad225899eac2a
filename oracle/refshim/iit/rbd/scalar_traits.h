// refshim/iit/rbd/scalar_traits.h — see rbd.h.  Trait classes the generated code calls its trigonometry through.
#pragma once
#include <cmath>
namespace iit {
namespace rbd {
struct DoubleTraits {
    typedef double Scalar;
    static double sin(const double& x) { return std::sin(x); }
    static double cos(const double& x) { return std::cos(x); }
    static double tan(const double& x) { return std::tan(x); }
    static double sqrt(const double& x) { return std::sqrt(x); }
    static double abs(const double& x) { return std::fabs(x); }
};
}  // namespace rbd
}  // namespace iit
