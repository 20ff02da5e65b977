// Minimal non-negative big integer for the HOST side of the engine: marshalling (bytes / limbs / decimal /
// hex), sampling, and the handful of per-key or per-proof scalar operations that are not worth a kernel
// launch (n*n, gcd, mod_inv for CRT constants).  It deliberately has no modular exponentiation: every
// BigInt::mod_pow of the reference's hot path runs in the CUDA kernels behind include/zkp_b200.h.
//
// Mirrors the slice of curv-kzen's `BigInt` API that reference/src/zkproofs/*.rs and src/serialize.rs use:
// to_bytes / from_bytes, to_str_radix / from_str_radix (10 and 16), bit_length, sample, sample_below,
// sample_range, div_floor, gcd, mod_inv, comparison and + - * %.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace zkhost {

class BigInt {
 public:
  std::vector<uint32_t> d;  // little-endian limbs, no trailing zeros (zero = empty)

  BigInt() {}
  BigInt(uint64_t v) {
    while (v) {
      d.push_back((uint32_t)v);
      v >>= 32;
    }
  }
  static BigInt zero() { return BigInt(); }
  static BigInt one() { return BigInt(1); }
  bool is_zero() const { return d.empty(); }
  bool is_odd() const { return !d.empty() && (d[0] & 1u); }

  void trim() {
    while (!d.empty() && d.back() == 0) d.pop_back();
  }
  size_t bit_length() const {
    if (d.empty()) return 0;
    return 32 * (d.size() - 1) + (32 - __builtin_clz(d.back()));
  }

  // ---- marshalling
  // BigInt::to_bytes (curv-kzen on GMP, RECALLED): minimal big-endian magnitude; zero -> {0x00}
  std::vector<uint8_t> to_bytes() const {
    size_t n = std::max<size_t>(1, (bit_length() + 7) / 8);
    std::vector<uint8_t> out(n, 0);
    for (size_t i = 0; i < n; ++i) {
      size_t limb = i / 4;
      if (limb < d.size()) out[n - 1 - i] = (uint8_t)(d[limb] >> (8 * (i % 4)));
    }
    return out;
  }
  static BigInt from_bytes(const uint8_t* p, size_t n) {
    BigInt r;
    r.d.assign((n + 3) / 4, 0);
    for (size_t i = 0; i < n; ++i) r.d[i / 4] |= (uint32_t)p[n - 1 - i] << (8 * (i % 4));
    r.trim();
    return r;
  }
  static BigInt from_bytes(const std::vector<uint8_t>& v) { return from_bytes(v.data(), v.size()); }
  // fixed-width little-endian limb row of the C ABI; throws if the value does not fit
  void to_limbs(uint32_t* out, size_t limbs) const {
    if (d.size() > limbs) throw std::length_error("BigInt wider than the limb row");
    std::fill(out, out + limbs, 0u);
    std::copy(d.begin(), d.end(), out);
  }
  std::vector<uint32_t> to_limbs(size_t limbs) const {
    std::vector<uint32_t> v(limbs);
    to_limbs(v.data(), limbs);
    return v;
  }
  static BigInt from_limbs(const uint32_t* p, size_t limbs) {
    BigInt r;
    r.d.assign(p, p + limbs);
    r.trim();
    return r;
  }
  // to_str_radix(10) / from_str_radix(s, 10)  (serialize.rs:10,24)
  std::string to_dec() const {
    if (d.empty()) return "0";
    std::vector<uint32_t> t(d);
    std::string s;
    while (!t.empty()) {
      uint64_t rem = 0;
      for (size_t i = t.size(); i-- > 0;) {
        uint64_t cur = (rem << 32) | t[i];
        t[i] = (uint32_t)(cur / 1000000000u);
        rem = cur % 1000000000u;
      }
      while (!t.empty() && t.back() == 0) t.pop_back();
      for (int k = 0; k < 9; ++k) {
        s.push_back((char)('0' + rem % 10));
        rem /= 10;
        if (t.empty() && rem == 0) break;
      }
    }
    while (s.size() > 1 && s.back() == '0') s.pop_back();
    std::reverse(s.begin(), s.end());
    return s;
  }
  static bool parse_dec(const std::string& s, BigInt& out) {
    if (s.empty()) return false;
    BigInt r;
    for (char ch : s) {
      if (ch < '0' || ch > '9') return false;
      uint64_t carry = (uint64_t)(ch - '0');
      for (size_t i = 0; i < r.d.size(); ++i) {
        uint64_t cur = (uint64_t)r.d[i] * 10u + carry;
        r.d[i] = (uint32_t)cur;
        carry = cur >> 32;
      }
      if (carry) r.d.push_back((uint32_t)carry);
    }
    r.trim();
    out = r;
    return true;
  }
  static BigInt from_dec(const std::string& s) {
    BigInt r;
    if (!parse_dec(s, r)) throw std::invalid_argument("invalid decimal BigInt");
    return r;
  }
  // curv-kzen's own serde for human-readable formats (RECALLED): lower-case hex of to_bytes()
  std::string to_hex_bytes() const {
    static const char* hx = "0123456789abcdef";
    std::string s;
    for (uint8_t b : to_bytes()) {
      s.push_back(hx[b >> 4]);
      s.push_back(hx[b & 15]);
    }
    return s;
  }
  static bool parse_hex_bytes(const std::string& s, BigInt& out) {
    if (s.size() % 2) return false;
    std::vector<uint8_t> v;
    auto nib = [](char c) -> int {
      if (c >= '0' && c <= '9') return c - '0';
      if (c >= 'a' && c <= 'f') return c - 'a' + 10;
      if (c >= 'A' && c <= 'F') return c - 'A' + 10;
      return -1;
    };
    for (size_t i = 0; i < s.size(); i += 2) {
      int a = nib(s[i]), b = nib(s[i + 1]);
      if (a < 0 || b < 0) return false;
      v.push_back((uint8_t)(a * 16 + b));
    }
    out = from_bytes(v);
    return true;
  }

  // ---- comparison
  static int cmp(const BigInt& a, const BigInt& b) {
    if (a.d.size() != b.d.size()) return a.d.size() < b.d.size() ? -1 : 1;
    for (size_t i = a.d.size(); i-- > 0;)
      if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
    return 0;
  }
  bool operator==(const BigInt& o) const { return cmp(*this, o) == 0; }
  bool operator!=(const BigInt& o) const { return cmp(*this, o) != 0; }
  bool operator<(const BigInt& o) const { return cmp(*this, o) < 0; }
  bool operator<=(const BigInt& o) const { return cmp(*this, o) <= 0; }
  bool operator>(const BigInt& o) const { return cmp(*this, o) > 0; }
  bool operator>=(const BigInt& o) const { return cmp(*this, o) >= 0; }

  // ---- arithmetic
  BigInt operator+(const BigInt& o) const {
    BigInt r;
    size_t n = std::max(d.size(), o.d.size());
    r.d.resize(n + 1);
    uint64_t c = 0;
    for (size_t i = 0; i < n; ++i) {
      uint64_t t = c + (i < d.size() ? d[i] : 0) + (uint64_t)(i < o.d.size() ? o.d[i] : 0);
      r.d[i] = (uint32_t)t;
      c = t >> 32;
    }
    r.d[n] = (uint32_t)c;
    r.trim();
    return r;
  }
  // requires *this >= o (the hot path never produces negatives; the reference would carry a sign)
  BigInt operator-(const BigInt& o) const {
    if (*this < o) throw std::domain_error("negative BigInt on the host path");
    BigInt r;
    r.d.resize(d.size());
    int64_t b = 0;
    for (size_t i = 0; i < d.size(); ++i) {
      int64_t t = (int64_t)d[i] - (i < o.d.size() ? o.d[i] : 0) - b;
      b = t < 0;
      r.d[i] = (uint32_t)(t + (b ? ((int64_t)1 << 32) : 0));
    }
    r.trim();
    return r;
  }
  BigInt operator*(const BigInt& o) const {
    BigInt r;
    if (d.empty() || o.d.empty()) return r;
    r.d.assign(d.size() + o.d.size(), 0);
    for (size_t i = 0; i < d.size(); ++i) {
      uint64_t c = 0;
      for (size_t j = 0; j < o.d.size(); ++j) {
        uint64_t t = (uint64_t)d[i] * o.d[j] + r.d[i + j] + c;
        r.d[i + j] = (uint32_t)t;
        c = t >> 32;
      }
      r.d[i + o.d.size()] = (uint32_t)c;
    }
    r.trim();
    return r;
  }
  BigInt operator|(const BigInt& o) const {
    BigInt r = d.size() >= o.d.size() ? *this : o;
    const BigInt& s = d.size() >= o.d.size() ? o : *this;
    for (size_t i = 0; i < s.d.size(); ++i) r.d[i] |= s.d[i];
    return r;
  }
  uint32_t mod_small(uint32_t m) const {  // this mod m for a 32-bit m
    uint64_t rem = 0;
    for (size_t i = d.size(); i-- > 0;) rem = ((rem << 32) | d[i]) % m;
    return (uint32_t)rem;
  }
  BigInt shl(size_t bits) const {
    if (d.empty()) return *this;
    BigInt r;
    size_t w = bits / 32, s = bits % 32;
    r.d.assign(d.size() + w + 1, 0);
    for (size_t i = 0; i < d.size(); ++i) {
      r.d[i + w] |= d[i] << s;
      if (s) r.d[i + w + 1] |= d[i] >> (32 - s);
    }
    r.trim();
    return r;
  }
  BigInt shr(size_t bits) const {
    size_t w = bits / 32, s = bits % 32;
    BigInt r;
    if (w >= d.size()) return r;
    r.d.assign(d.size() - w, 0);
    for (size_t i = 0; i < r.d.size(); ++i) {
      r.d[i] = d[i + w] >> s;
      if (s && i + w + 1 < d.size()) r.d[i] |= d[i + w + 1] << (32 - s);
    }
    r.trim();
    return r;
  }
  // Knuth algorithm D.  q = floor(a / b), r = a mod b.
  static void divmod(const BigInt& a, const BigInt& b, BigInt& q, BigInt& r) {
    if (b.d.empty()) throw std::domain_error("division by zero");
    if (cmp(a, b) < 0) {
      q = BigInt();
      r = a;
      return;
    }
    if (b.d.size() == 1) {
      q.d.assign(a.d.size(), 0);
      uint64_t rem = 0;
      for (size_t i = a.d.size(); i-- > 0;) {
        uint64_t cur = (rem << 32) | a.d[i];
        q.d[i] = (uint32_t)(cur / b.d[0]);
        rem = cur % b.d[0];
      }
      q.trim();
      r = BigInt(rem);
      return;
    }
    int s = __builtin_clz(b.d.back());
    BigInt v = b.shl(s), u = a.shl(s);
    size_t n = v.d.size(), m = u.d.size() >= n ? u.d.size() - n : 0;
    u.d.resize(n + m + 1, 0);
    q.d.assign(m + 1, 0);
    for (size_t j = m + 1; j-- > 0;) {
      uint64_t num = ((uint64_t)u.d[j + n] << 32) | u.d[j + n - 1];
      uint64_t qhat = num / v.d[n - 1], rhat = num % v.d[n - 1];
      while (qhat >= ((uint64_t)1 << 32) || qhat * v.d[n - 2] > ((rhat << 32) | u.d[j + n - 2])) {
        --qhat;
        rhat += v.d[n - 1];
        if (rhat >= ((uint64_t)1 << 32)) break;
      }
      int64_t borrow = 0;
      uint64_t carry = 0;
      for (size_t i = 0; i < n; ++i) {
        uint64_t p = qhat * v.d[i] + carry;
        carry = p >> 32;
        int64_t t = (int64_t)u.d[i + j] - borrow - (int64_t)(p & 0xffffffffu);
        borrow = t < 0;
        u.d[i + j] = (uint32_t)t;
      }
      int64_t t = (int64_t)u.d[j + n] - borrow - (int64_t)carry;
      borrow = t < 0;
      u.d[j + n] = (uint32_t)t;
      if (borrow) {
        --qhat;
        uint64_t c = 0;
        for (size_t i = 0; i < n; ++i) {
          uint64_t s2 = (uint64_t)u.d[i + j] + v.d[i] + c;
          u.d[i + j] = (uint32_t)s2;
          c = s2 >> 32;
        }
        u.d[j + n] += (uint32_t)c;
      }
      q.d[j] = (uint32_t)qhat;
    }
    q.trim();
    u.d.resize(n);
    u.trim();
    r = u.shr(s);
  }
  BigInt operator/(const BigInt& o) const {
    BigInt q, r;
    divmod(*this, o, q, r);
    return q;
  }
  BigInt operator%(const BigInt& o) const {
    BigInt q, r;
    divmod(*this, o, q, r);
    return r;
  }
  BigInt div_floor(const BigInt& o) const { return *this / o; }  // operands are non-negative here

  static BigInt gcd(BigInt a, BigInt b) {
    while (!b.is_zero()) {
      BigInt r = a % b;
      a = b;
      b = r;
    }
    return a;
  }
  // BigInt::mod_inv(a, m) -> Option: returns false when gcd(a, m) != 1
  static bool mod_inv(const BigInt& a, const BigInt& m, BigInt& out) {
    // extended Euclid on (r0, r1) with coefficients tracked modulo m as (value, negative?) pairs
    BigInt r0 = m, r1 = a % m, t0 = BigInt(), t1 = BigInt(1);
    bool n0 = false, n1 = false;
    while (!r1.is_zero()) {
      BigInt q, r2;
      divmod(r0, r1, q, r2);
      // t2 = t0 - q*t1
      BigInt qt = q * t1;
      BigInt t2;
      bool n2;
      if (n0 == n1) {
        if (t0 >= qt) { t2 = t0 - qt; n2 = n0; } else { t2 = qt - t0; n2 = !n0; }
      } else {
        t2 = t0 + qt;
        n2 = n0;
      }
      r0 = r1; r1 = r2;
      t0 = t1; n0 = n1;
      t1 = t2; n1 = n2;
    }
    if (r0 != BigInt(1)) return false;
    BigInt t = t0 % m;
    out = (n0 && !t.is_zero()) ? m - t : t;
    return true;
  }

  // ---- sampling (curv-kzen, RECALLED).  `fill(buf, n)` supplies RNG bytes.
  using ByteSource = std::function<void(uint8_t*, size_t)>;
  static BigInt sample(const ByteSource& fill, size_t bits) {
    size_t nbytes = (bits + 7) / 8;
    std::vector<uint8_t> buf(nbytes);
    fill(buf.data(), nbytes);
    return from_bytes(buf).shr(8 * nbytes - bits);
  }
  static BigInt sample_below(const ByteSource& fill, const BigInt& upper) {
    size_t bits = upper.bit_length();
    for (;;) {
      BigInt v = sample(fill, bits);
      if (v < upper) return v;
    }
  }
  static BigInt sample_range(const ByteSource& fill, const BigInt& lo, const BigInt& hi) { return lo + sample_below(fill, hi - lo); }
};

}  // namespace zkhost
