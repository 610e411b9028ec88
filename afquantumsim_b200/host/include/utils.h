/* utils.h — host bit/string helpers of the afQuantumSim API (reference
 * include/utils.h).  The array helpers of the reference (tensor_product,
 * insert_bits, gen_index; src/utils.cpp:137-195) have no counterpart here: the
 * engine never builds embedded operator matrices. */
#pragma once
#include <arrayfire.h>

#include <cstdint>
#include <string>
#include <utility>

/* NB: like the reference, these live in the global namespace. */

/* str repeated n times (src/utils.cpp:32-40) */
std::string repeat(std::size_t n, const std::string& str);
/* number of UTF-8 code points */
std::size_t utf8str_len(std::string str);
/* `length` binary digits of val, most significant first */
std::string binary_string(uint32_t val, int length);
/* reverse the low `length` bits of val (src/utils.cpp:42-53) */
uint32_t reverse_binary(uint32_t val, int length) noexcept;
/* bits [first, last] of val, counted from the LSB */
uint32_t extract_binary(uint32_t val, int first, int last) noexcept;
/* continued-fraction approximation num/den of value with den <= max_denominator */
std::pair<int64_t, int64_t> approximate_fraction(double value, int64_t max_denominator);
int64_t gcd(int64_t a, int64_t b);

static inline uint32_t fast_pow2(uint32_t pow) { return 1u << pow; }
static inline uint32_t fast_log2(uint32_t val) {
    int counter = -(val != 0);
    for (; val; counter++) val >>= 1;
    return counter;
}

/* Kronecker product on host arrays (reference src/utils.cpp:16-30); kept for API
 * compatibility, small inputs only. */
af::array tensor_product(const af::array& lhs, const af::array& rhs);
