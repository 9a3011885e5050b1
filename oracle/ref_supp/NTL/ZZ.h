/* oracle/ref_supp/NTL/ZZ.h -- TEST INFRASTRUCTURE. Header-only stand-in exposing exactly the NTL
 * surface mlwe_verifier.cpp touches: ZZ(long), ZZ_p::init, ZZ_p(long), assignment from
 * integers, operator!=, Vec<T>::SetLength/operator[], ZZ_pX, interpolate, eval, conv<int>.
 * Semantics per ntl/share/doc/NTL/ZZ_pX.txt:408-419 (unique interpolant of degree < n, evaluation);
 * results are canonical residues, hence implementation-independent. */
#ifndef KOSK_NTL_SHIM_H
#define KOSK_NTL_SHIM_H
#include <vector>
#include <cstddef>
namespace NTL {
struct ZZ { long v; ZZ() : v(0) {} explicit ZZ(long x) : v(x) {} };
struct ZZ_p {
    long v;
    static long &modulus() { static long m = 3329; return m; }
    static void init(const ZZ &m) { modulus() = m.v; }
    ZZ_p() : v(0) {}
    explicit ZZ_p(long x) : v(((x % modulus()) + modulus()) % modulus()) {}
    ZZ_p &operator=(long x) { v = ((x % modulus()) + modulus()) % modulus(); return *this; }
};
inline bool operator!=(const ZZ_p &a, long b) { return a.v != ((b % ZZ_p::modulus()) + ZZ_p::modulus()) % ZZ_p::modulus(); }
inline bool operator==(const ZZ_p &a, long b) { return !(a != b); }
template <class T> struct Vec {
    std::vector<T> d;
    void SetLength(long n) { d.resize((size_t)n); }
    long length() const { return (long)d.size(); }
    T &operator[](long i) { return d[(size_t)i]; }
    const T &operator[](long i) const { return d[(size_t)i]; }
};
struct ZZ_pX { std::vector<long> c; };
inline long kosk_powm(long b, long e, long q) { long r = 1; b %= q; while (e) { if (e & 1) r = r * b % q; b = b * b % q; e >>= 1; } return r; }
/* Newton divided differences -> monomial coefficients, O(n^2) */
inline void interpolate(ZZ_pX &f, const Vec<ZZ_p> &a, const Vec<ZZ_p> &b)
{
    const long q = ZZ_p::modulus(); const long n = a.length();
    std::vector<long> dd(n), x(n);
    for (long i = 0; i < n; i++) { dd[i] = b[i].v; x[i] = a[i].v; }
    for (long j = 1; j < n; j++)
        for (long i = n - 1; i >= j; i--) {
            long den = ((x[i] - x[i - j]) % q + q) % q;
            dd[i] = ((dd[i] - dd[i - 1]) % q + q) % q * kosk_powm(den, q - 2, q) % q;
        }
    f.c.assign((size_t)n, 0);
    /* Horner over Newton basis: p = dd[n-1]; p = p*(X - x[i]) + dd[i] */
    std::vector<long> p(1, n ? dd[n - 1] : 0);
    for (long i = n - 2; i >= 0; i--) {
        std::vector<long> np(p.size() + 1, 0);
        for (size_t k = 0; k < p.size(); k++) {
            np[k + 1] = (np[k + 1] + p[k]) % q;
            np[k] = ((np[k] - p[k] * x[i]) % q + q) % q;
        }
        np[0] = (np[0] + dd[i]) % q;
        p.swap(np);
    }
    for (size_t k = 0; k < p.size(); k++) f.c[k] = p[k];
}
inline ZZ_p eval(const ZZ_pX &f, const ZZ_p &a)
{
    const long q = ZZ_p::modulus(); long r = 0;
    for (size_t k = f.c.size(); k-- > 0;) r = (r * a.v + f.c[k]) % q;
    ZZ_p o; o.v = r; return o;
}
template <class T> inline T conv(const ZZ_p &a) { return (T)a.v; }
}
#endif
