/*
 * arrayfire.h — compatibility shim, NOT ArrayFire.
 *
 * afQuantumSim's public headers expose a few ArrayFire types (af::cfloat,
 * af::Backend, af::array in QCircuit::circuit(), QSimulator::statevector(),
 * QSimulator(n, af::array), QState::to_array(); include/quantum.h:292,403,532,726).
 * ArrayFire is not part of this build: all arithmetic runs in the CUDA engine
 * behind include/aqs_engine.h.  This header provides just enough of those types,
 * as plain HOST containers, for client code written against the reference API
 * (its tests, examples and benchmark) to compile unchanged.  An af::array here
 * is a host-side snapshot; it owns no device memory and performs no device math.
 */
#pragma once

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace af {

struct cfloat {
    float real;
    float imag;
    cfloat() : real(0.f), imag(0.f) {}
    cfloat(float r) : real(r), imag(0.f) {}
    template<typename A, typename B,
             typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
    cfloat(A r, B i) : real(static_cast<float>(r)), imag(static_cast<float>(i)) {}
    cfloat(const std::complex<float>& z) : real(z.real()), imag(z.imag()) {}
    operator std::complex<float>() const { return {real, imag}; }
};

inline cfloat operator+(const cfloat& a, const cfloat& b) { return {a.real + b.real, a.imag + b.imag}; }
inline cfloat operator-(const cfloat& a, const cfloat& b) { return {a.real - b.real, a.imag - b.imag}; }
inline cfloat operator-(const cfloat& a) { return {-a.real, -a.imag}; }
inline cfloat operator*(const cfloat& a, const cfloat& b) {
    return {a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real};
}
inline cfloat operator*(const cfloat& a, float b) { return {a.real * b, a.imag * b}; }
inline cfloat operator*(float a, const cfloat& b) { return {a * b.real, a * b.imag}; }
inline cfloat operator/(const cfloat& a, const cfloat& b) {
    const float d = b.real * b.real + b.imag * b.imag;
    return {(a.real * b.real + a.imag * b.imag) / d, (a.imag * b.real - a.real * b.imag) / d};
}
/* A real divisor is promoted to complex and divided with the textbook formula
 * (x*c/(c*c)), not x/c: that is the rounding under which the reference's exact
 * comparison of a CH result with a normalised product state holds
 * (test/tests.cpp:986-1002; see DESIGN.md "QState normalisation"). */
inline cfloat operator/(const cfloat& a, float b) { return a / cfloat(b, 0.f); }
inline bool operator==(const cfloat& a, const cfloat& b) { return a.real == b.real && a.imag == b.imag; }
inline bool operator!=(const cfloat& a, const cfloat& b) { return !(a == b); }
inline float real(const cfloat& a) { return a.real; }
inline float imag(const cfloat& a) { return a.imag; }
inline float abs(const cfloat& a) { return std::sqrt(a.real * a.real + a.imag * a.imag); }
inline cfloat conj(const cfloat& a) { return {a.real, -a.imag}; }
template<typename OS>
OS& operator<<(OS& os, const cfloat& z) {
    os << "(" << z.real << "," << z.imag << ")";
    return os;
}

enum Backend {
    AF_BACKEND_DEFAULT = 0,
    AF_BACKEND_CPU     = 1,
    AF_BACKEND_CUDA    = 2,
    AF_BACKEND_OPENCL  = 4
};
enum dtype { f32, c32, b8, u32, s32 };

struct dim4 {
    long long d[4];
    dim4(long long a = 1, long long b = 1, long long c = 1, long long e = 1) : d{a, b, c, e} {}
    long long operator[](int i) const { return d[i]; }
    long long& operator[](int i) { return d[i]; }
    long long elements() const { return d[0] * d[1] * d[2] * d[3]; }
};

class exception : public std::runtime_error {
   public:
    using std::runtime_error::runtime_error;
};

/* Host-side column-major array of complex64 (real-valued and boolean results
 * keep their value in .real). */
class array {
   public:
    array() : dims_(0, 1, 1, 1), type_(c32) {}
    array(long long n, dtype t = c32) : dims_(n), type_(t), data_(static_cast<size_t>(n)) {}
    array(long long r, long long c, dtype t) : dims_(r, c), type_(t), data_(static_cast<size_t>(r * c)) {}
    array(long long n, const cfloat* host) : dims_(n), type_(c32), data_(host, host + n) {}
    array(long long r, long long c, const cfloat* host) : dims_(r, c), type_(c32), data_(host, host + r * c) {}

    const dim4& dims() const { return dims_; }
    long long dims(int i) const { return dims_[i]; }
    long long elements() const { return static_cast<long long>(data_.size()); }
    bool isempty() const { return data_.empty(); }
    void eval() const {}   /* ArrayFire forces lazy evaluation here; snapshots are already concrete */
    dtype type() const { return type_; }

    void host(void* out) const {
        if (type_ == c32) {
            std::copy(data_.begin(), data_.end(), static_cast<cfloat*>(out));
        } else {
            float* f = static_cast<float*>(out);
            for (size_t i = 0; i < data_.size(); ++i) f[i] = data_[i].real;
        }
    }
    template<typename T>
    T* host() const {
        T* p = new T[data_.size()];
        host(p);
        return p;
    }
    cfloat* data() { return data_.data(); }
    const cfloat* data() const { return data_.data(); }
    std::vector<cfloat>& storage() { return data_; }
    const std::vector<cfloat>& storage() const { return data_; }

    struct elem {
        cfloat v;
        template<typename T>
        T scalar() const;
    };
    elem operator()(long long i) const { return elem{data_.at(static_cast<size_t>(i))}; }
    elem operator()(long long r, long long c) const { return elem{data_.at(static_cast<size_t>(c * dims_[0] + r))}; }
    template<typename T>
    T scalar() const {
        return elem{data_.at(0)}.template scalar<T>();
    }

    array T() const {
        array out(dims_[1], dims_[0], type_);
        for (long long c = 0; c < dims_[1]; ++c)
            for (long long r = 0; r < dims_[0]; ++r) out.data_[static_cast<size_t>(r * dims_[1] + c)] = data_[static_cast<size_t>(c * dims_[0] + r)];
        return out;
    }

   private:
    dim4 dims_;
    dtype type_;
    std::vector<cfloat> data_;
    friend array elementwise(const array&, const array&, int);
};

template<>
inline cfloat array::elem::scalar<cfloat>() const { return v; }
template<>
inline float array::elem::scalar<float>() const { return v.real; }
template<>
inline uint32_t array::elem::scalar<uint32_t>() const { return static_cast<uint32_t>(v.real); }
template<>
inline bool array::elem::scalar<bool>() const { return v.real != 0.f; }

inline void check_same_shape(const array& a, const array& b) {
    if (a.elements() != b.elements()) throw exception("af shim: size mismatch");
}
inline array operator==(const array& a, const array& b) {
    check_same_shape(a, b);
    array out(a.dims(0), a.dims(1), b8);
    for (long long i = 0; i < a.elements(); ++i) out.data()[i] = cfloat(a.data()[i] == b.data()[i] ? 1.f : 0.f);
    return out;
}
inline array operator-(const array& a, const array& b) {
    check_same_shape(a, b);
    array out(a.dims(0), a.dims(1), a.type());
    for (long long i = 0; i < a.elements(); ++i) out.data()[i] = a.data()[i] - b.data()[i];
    return out;
}
inline array abs(const array& a) {
    array out(a.dims(0), a.dims(1), f32);
    for (long long i = 0; i < a.elements(); ++i) out.data()[i] = cfloat(abs(a.data()[i]));
    return out;
}
inline array operator<(const array& a, double v) {
    array out(a.dims(0), a.dims(1), b8);
    for (long long i = 0; i < a.elements(); ++i) out.data()[i] = cfloat(a.data()[i].real < v ? 1.f : 0.f);
    return out;
}
template<typename T>
T allTrue(const array& a) {
    for (long long i = 0; i < a.elements(); ++i)
        if (a.data()[i].real == 0.f && a.data()[i].imag == 0.f) return static_cast<T>(false);
    return static_cast<T>(true);
}
inline array transpose(const array& a, bool conjugate = false) {
    array out = a.T();
    if (conjugate)
        for (long long i = 0; i < out.elements(); ++i) out.data()[i] = conj(out.data()[i]);
    return out;
}

/* implemented in the host library: names the CUDA device the engine runs on */
std::string infoString();
void info();
void sync();
void print(const char* name, const array& a);

}  // namespace af

#define af_print(x) af::print(#x, x)
