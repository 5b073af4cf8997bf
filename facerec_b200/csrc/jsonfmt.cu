// Host-side writer helper: the 512-float embedding of a features.jsonl line as JSON text, byte-identical
// to Python's json.dumps(list_of_floats, separators=(",", ":")) for float32 values widened to double (what
// `embedding.tolist()` yields in facerec/extract.py:166,186-195).  Python's float repr is the shortest string that
// round-trips the double, laid out in fixed notation for 1e-4 <= |x| < 1e16 and as d.ddde-XX otherwise;
// std::to_chars gives the same shortest digits, the layout rules are applied here.
#include <charconv>
#include <cmath>
#include <cstring>

#include "../../include/facerec_b200.h"

namespace {

// writes repr(x) at p (at most 32 bytes), returns the end
char* py_float_repr(double x, char* p) {
  if (std::isnan(x)) { memcpy(p, "NaN", 3); return p + 3; }                 // json.dumps(allow_nan=True) spellings
  if (std::isinf(x)) {
    if (x < 0) *p++ = '-';
    memcpy(p, "Infinity", 8);
    return p + 8;
  }
  char sci[40];
  auto r = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific);
  const char* s = sci;
  const char* end = r.ptr;
  if (*s == '-') { *p++ = '-'; ++s; }
  const char* epos = s;
  while (epos < end && *epos != 'e') ++epos;
  char digits[24];
  int nd = 0;
  for (const char* q = s; q < epos; ++q)
    if (*q != '.') digits[nd++] = *q;
  int e = 0;
  {
    const char* q = epos + 1;
    bool neg = false;
    if (*q == '-') { neg = true; ++q; } else if (*q == '+') ++q;
    for (; q < end; ++q) e = e * 10 + (*q - '0');
    if (neg) e = -e;
  }
  if (nd == 1 && digits[0] == '0') { memcpy(p, "0.0", 3); return p + 3; }
  if (e >= -4 && e < 16) {
    if (e >= 0) {
      for (int i = 0; i <= e; ++i) *p++ = i < nd ? digits[i] : '0';
      *p++ = '.';
      if (nd > e + 1) {
        memcpy(p, digits + e + 1, nd - e - 1);
        p += nd - e - 1;
      } else {
        *p++ = '0';
      }
    } else {
      *p++ = '0';
      *p++ = '.';
      for (int i = 0; i < -e - 1; ++i) *p++ = '0';
      memcpy(p, digits, nd);
      p += nd;
    }
    return p;
  }
  *p++ = digits[0];
  if (nd > 1) {
    *p++ = '.';
    memcpy(p, digits + 1, nd - 1);
    p += nd - 1;
  }
  *p++ = 'e';
  *p++ = e < 0 ? '-' : '+';
  int a = e < 0 ? -e : e;
  char tmp[8];
  int nt = 0;
  do { tmp[nt++] = static_cast<char>('0' + a % 10); a /= 10; } while (a);
  if (nt < 2) tmp[nt++] = '0';
  while (nt) *p++ = tmp[--nt];
  return p;
}

}  // namespace

extern "C" long long frb_format_f32_json(const float* v, int n, char* out, long long cap) {
  if (v == nullptr || out == nullptr || n < 0 || cap < 2 + 33LL * n) return -1;
  char* p = out;
  *p++ = '[';
  for (int i = 0; i < n; ++i) {
    if (i) *p++ = ',';
    p = py_float_repr(static_cast<double>(v[i]), p);
  }
  *p++ = ']';
  return p - out;
}
