// types/VecX.h — minimal dynamic column vector standing in for Eigen::Matrix<Scalar, Dynamic, 1>.
// Eigen is not part of this build; the reference's System / Trainer interfaces only need size(), operator[],
// Zero(n), Ones(n), +=, and division by a scalar (reference src/types/System.h:11-12, src/Trainer.cpp:61-63,192).
#pragma once
#include <cstddef>
#include <vector>

namespace avs {

template <class S> class VecX {
public:
    VecX() {}
    explicit VecX(std::size_t n) : v_(n) {}
    VecX(std::size_t n, const S& fill) : v_(n, fill) {}
    static VecX Zero(std::size_t n) { return VecX(n, S(0)); }
    static VecX Ones(std::size_t n) { return VecX(n, S(1)); }
    std::size_t size() const { return v_.size(); }
    void resize(std::size_t n) { v_.resize(n); }
    S& operator[](std::size_t i) { return v_[i]; }
    const S& operator[](std::size_t i) const { return v_[i]; }
    S& operator()(std::size_t i) { return v_[i]; }
    const S& operator()(std::size_t i) const { return v_[i]; }
    S* data() { return v_.data(); }
    const S* data() const { return v_.data(); }
    VecX& operator+=(const VecX& o) { for (std::size_t i = 0; i < v_.size(); i++) v_[i] += o.v_[i]; return *this; }
    VecX operator/(double d) const { VecX r(*this); for (auto& x : r.v_) x = x / d; return r; }
    VecX operator*(double d) const { VecX r(*this); for (auto& x : r.v_) x = x * d; return r; }
private:
    std::vector<S> v_;
};

}  // namespace avs

// source compatibility for client code that spells the vector type through Eigen
namespace Eigen {
const int Dynamic = -1;
template <class S, int Rows, int Cols = 1> using Matrix = avs::VecX<S>;
}  // namespace Eigen
