// hostmath.hpp -- host-side number theory for table generation (native C++
// mirror of the reference's precompute: ring/ring_context.go:68-209,
// ring/utils.go:25-288, ring/ring_basis_extension.go:39-142).  Runs once per
// context; everything it produces is uploaded to the device.
#pragma once
#include <stdint.h>

#include <vector>

namespace lgh {

typedef uint64_t u64;
typedef unsigned __int128 u128;

inline u64 mulmod(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }
inline u64 powmod(u64 a, u64 e, u64 m) {
    u64 r = 1 % m;
    a %= m;
    for (; e; e >>= 1) {
        if (e & 1) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
    }
    return r;
}
inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

// BRedParams (modular_reduction.go:97-106): {hi, lo} of floor(2^128 / q)
inline void bred_params(u64 q, u64& hi, u64& lo) {
    const u128 b = (u128)1 << 64;
    hi = (u64)(b / q);
    lo = (u64)(((u128)(u64)(b % q) << 64) / q);
}
// MRedParams (:53-64): q^-1 mod 2^64 by Newton iteration (same value as q^(2^63-1))
inline u64 mred_params(u64 q) {
    u64 x = q;  // correct to 3 bits
    for (int i = 0; i < 6; ++i) x *= 2 - q * x;
    return x;
}
// MForm (:15-22): a * 2^64 mod q for a < q
inline u64 mform(u64 a, u64 q) { return (u64)((((u128)a) << 64) % q); }
inline u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    const u128 p = (u128)x * y;
    const u64 H = mulhi((u64)p * qinv, q);
    u64 r = (u64)(p >> 64) - H + q;
    return r >= q ? r - q : r;
}
inline u64 bitrev(u64 x, unsigned bits) {
    u64 r = 0;
    for (unsigned i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
inline unsigned log2u(u64 n) {
    unsigned l = 0;
    while (((u64)1 << l) < n) ++l;
    return l;
}

inline const std::vector<u64>& small_primes() {  // ring/utils.go:290-391 = primes < 17390
    static std::vector<u64> p;
    if (p.empty()) {
        std::vector<char> comp(17390, 0);
        for (int i = 2; i < 17390; ++i)
            if (!comp[i]) {
                p.push_back(i);
                for (int j = i * i; j < 17390; j += i) comp[j] = 1;
            }
    }
    return p;
}

// IsPrime (ring/utils.go:75-128) with deterministic Miller-Rabin witnesses
inline bool is_prime(u64 n) {
    if (n < 2) return false;
    for (u64 p : small_primes()) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    u64 d = n - 1;
    int k = 0;
    while (!(d & 1)) {
        d >>= 1;
        ++k;
    }
    static const u64 wit[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 a : wit) {
        u64 x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < k; ++i) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}

inline u64 gcd(u64 a, u64 b) {
    if (!a || !b) return 0;
    while (b) {
        u64 t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// The reference picks psi from the smallest g >= 3 that passes its test against
// the factor list produced by trial division + its Pollard-rho walk
// (ring/utils.go:182-288).  The walk (x -> x^2 + c, c = 1..9, start 2, ordered
// pair before the gcd) is reproduced so that the factor list -- and therefore
// g, psi and the NTT output ordering -- is the reference's.
inline u64 rho_step(u64 x, u64 m, u64 c) { return (mulmod(x % m, x % m, m) + c) % m; }
inline u64 rho_factor(u64 m) {
    u64 d = 0;
    for (u64 c = 1; c < 10; ++c) {
        u64 x = 2, y = 2;
        d = 1;
        while (d != 0) {
            x = rho_step(x, m, c);
            y = rho_step(rho_step(y, m, c), m, c);
            if (y > x) {
                u64 t = x;
                x = y;
                y = t;
            }
            d = gcd(x - y, m);
            if (d > 1) return d;
        }
    }
    return d;
}
inline std::vector<u64> factors_of(u64 n) {
    std::vector<u64> f;
    u64 m = n;
    for (u64 p : small_primes()) {
        bool hit = false;
        while (m % p == 0) {
            m /= p;
            hit = true;
        }
        if (hit) f.push_back(p);
    }
    if (m == 1) return f;
    for (;;) {
        u64 d = rho_factor(m);
        if (d == 0) {
            f.push_back(m);
            break;
        }
        m /= d;
        if (!f.empty() && d == f.back()) continue;
        f.push_back(d);
    }
    return f;
}
inline u64 primitive_root(u64 q) {
    const std::vector<u64> f = factors_of(q - 1);
    for (u64 g = 3;; ++g) {
        bool ok = true;
        for (u64 p : f)
            if (powmod(g, (q - 1) / p, q) == 1) {
                ok = false;
                break;
            }
        if (ok) return g;
    }
}

}  // namespace lgh
