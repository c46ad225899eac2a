// refshim/iit/rbd/TransformsBase.h — see rbd.h.  Matrices that are functions of the joint state: calling the object with a
// state re-evaluates the state-dependent entries through the derived class' update() (CRTP), as in iit-rbd.
#pragma once
#include "rbd.h"
namespace iit {
namespace rbd {
template <class State, int Rows, int Cols, class ActualMatrix>
class StateDependentMatrix : public PlainMatrix<typename State::Scalar, Rows, Cols> {
public:
    typedef typename State::Scalar Scalar;
    typedef PlainMatrix<Scalar, Rows, Cols> Base;
    typedef Base MatrixType;
    StateDependentMatrix() { this->setZero(); }
    using Base::operator();
    const ActualMatrix& operator()(State const& s) { return static_cast<ActualMatrix*>(this)->update(s); }
};
template <class State, class ActualMatrix> class HomogeneousTransformBase : public StateDependentMatrix<State, 4, 4, ActualMatrix> {};
template <class State, class ActualMatrix> class RotationTransformBase : public StateDependentMatrix<State, 3, 3, ActualMatrix> {};
template <class State, class ActualMatrix> class SpatialTransformBase : public StateDependentMatrix<State, 6, 6, ActualMatrix> {};
template <class State, int Cols, class ActualMatrix> class JacobianBase : public StateDependentMatrix<State, 6, Cols, ActualMatrix> {};
}  // namespace rbd
}  // namespace iit
