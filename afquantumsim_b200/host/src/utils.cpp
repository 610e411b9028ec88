// utils.cpp — host helpers (global namespace, like the reference's src/utils.cpp).
#include "utils.h"

#include <cmath>
#include <stdexcept>

std::string repeat(std::size_t n, const std::string& str) {
    std::string out;
    out.reserve(n * str.size());
    for (std::size_t i = 0; i < n; ++i) out += str;
    return out;
}

std::size_t utf8str_len(std::string str) {
    std::size_t len = 0;
    for (unsigned char c : str)
        if ((c & 0xC0) != 0x80) ++len;   // count lead bytes only
    return len;
}

std::string binary_string(uint32_t val, int length) {
    std::string out(static_cast<std::size_t>(length), '0');
    for (int i = 0; i < length; ++i)
        if (val & (1u << (length - i - 1))) out[static_cast<std::size_t>(i)] = '1';
    return out;
}

uint32_t reverse_binary(uint32_t val, int length) noexcept {
    uint32_t out = length >= 32 ? 0u : (val & ~((1u << length) - 1u));   // bits above `length` are kept
    for (int i = 0; i < length; ++i)
        if (val & (1u << i)) out |= 1u << (length - 1 - i);
    return out;
}

uint32_t extract_binary(uint32_t val, int first, int last) noexcept {
    return (val >> first) & ((1u << (1 + last - first)) - 1u);
}

int64_t gcd(int64_t a, int64_t b) {
    while (b != 0) {
        int64_t t = a % b;
        a         = b;
        b         = t;
    }
    return a;
}

// Best rational approximation with bounded denominator, by walking the
// continued-fraction convergents of the (exactly represented) binary value and
// finishing with the best semiconvergent.
std::pair<int64_t, int64_t> approximate_fraction(double value, int64_t max_denominator) {
    if (max_denominator <= 1) return {static_cast<int64_t>(value), 1};
    const bool negative = value < 0;
    if (negative) value = -value;

    int64_t den = 1;
    while (value != std::floor(value)) {   // value = num / den exactly (den a power of two)
        den <<= 1;
        value *= 2;
    }
    int64_t num = static_cast<int64_t>(value);

    int64_t h0 = 0, h1 = 1, k0 = 1, k1 = 0;   // convergents h/k
    for (int i = 0; i < 64; ++i) {
        const int64_t a = den ? num / den : 0;
        if (i && !a) break;
        const int64_t rem = den ? num % den : 0;
        num               = den;
        den               = rem;

        int64_t term = a;
        bool stop    = false;
        if (k1 * a + k0 >= max_denominator) {
            term = (max_denominator - k0) / k1;
            if (term * 2 >= a || k1 >= max_denominator) stop = true;   // semiconvergent is the better one
            else break;
        }
        const int64_t h2 = term * h1 + h0, k2 = term * k1 + k0;
        h0 = h1; h1 = h2;
        k0 = k1; k1 = k2;
        if (stop) break;
    }
    return {negative ? -h1 : h1, k1};
}

af::array tensor_product(const af::array& lhs, const af::array& rhs) {
    if (lhs.dims(2) != 1 || lhs.dims(3) != 1 || rhs.dims(2) != 1 || rhs.dims(3) != 1)
        throw std::invalid_argument{"Cannot compute tensor product of arrays with more than 2 dimensions"};
    const long long lr = lhs.dims(0), lc = lhs.dims(1), rr = rhs.dims(0), rc = rhs.dims(1);
    af::array out(lr * rr, lc * rc, af::c32);
    for (long long c = 0; c < lc * rc; ++c)
        for (long long r = 0; r < lr * rr; ++r)
            out.data()[c * (lr * rr) + r] =
                lhs.data()[(c / rc) * lr + (r / rr)] * rhs.data()[(c % rc) * rr + (r % rr)];
    return out;
}
