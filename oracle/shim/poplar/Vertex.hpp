// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Minimal stand-in for <poplar/Vertex.hpp> so that the reference's codelet
// sources (/root/reference/ba/gbp_codelets.cpp + matlib.cpp + bafuncs.cpp)
// compile UNMODIFIED with g++ into oracle/_ref/ (see oracle/Makefile).  It
// provides only what those three files use: field
// wrappers that are bound to raw host pointers before compute() is called.
#pragma once
#include <cstddef>

namespace poplar {

struct Vertex {};

template <class T>
struct Vector {
  T* ptr = nullptr;
  std::size_t len = 0;
  T& operator[](std::size_t i) const { return ptr[i]; }
  std::size_t size() const { return len; }
  void bind(T* p, std::size_t n) {
    ptr = p;
    len = n;
  }
};

template <class T>
struct ScalarField {
  T* ptr = nullptr;
  void bind(T* p) { ptr = p; }
  operator const T&() const { return *ptr; }
};

template <class T>
struct Input : ScalarField<T> {};

template <class T>
struct InOut : ScalarField<T> {
  T& operator*() const { return *this->ptr; }
  InOut& operator-=(const T& v) {
    *this->ptr -= v;
    return *this;
  }
};

template <class T>
struct Output : ScalarField<T> {
  T& operator*() const { return *this->ptr; }
};

template <class T>
struct Input<Vector<T>> : Vector<T> {};
template <class T>
struct InOut<Vector<T>> : Vector<T> {};
template <class T>
struct Output<Vector<T>> : Vector<T> {};

}  // namespace poplar
