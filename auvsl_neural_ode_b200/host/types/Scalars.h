// types/Scalars.h — scalar type of the drop-in host API.
// The reference runs on CppAD::AD<double> (reference src/types/Scalars.h:11) because CppAD's tape is its
// gradient engine.  Here the gradient comes from the hand-written adjoint kernel behind the C ABI, so the
// host-visible scalar is a plain double; the name ADF is kept so that client code compiles unchanged.
#pragma once
typedef double ADF;
