// fortran_fmt.hpp -- the number formats of the reference's formatted writes, and a reader for them.
//
// The reference writes with WIDTH-LESS edit descriptors: '(e)', '(f)', '(es)' (a DEC/Intel extension; gfortran needs
// -fdec-format-defaults).  For real(8) both compilers then use w = 25, d = 16 and a two-digit exponent:
// E25.16 -> "   0.1000000000000000E-01", F25.16 -> "   22087.0000000000000000", ES25.16 -> "   1.0000000000000000E-02".
// Files shipped with the reference that were written this way pin it: INPUT_DOS/Al2O3.dos ("  -0.7400000000000000E+01"),
// INPUT_CDF/Ru.cdf ("22087.0000000000000000", "0.6095683453463127E+00").
// The reference READS its table cache back with the same descriptors and advance='no' (Analytical_IMFPs.f90:602-634,
// Reading_files_and_parameters.f90:2875-2900), i.e. as fixed fields of 25 characters: the widths matter, not only the values.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace trk3 {

// Fortran Fw.d: right-justified, asterisks on overflow
inline std::string fmt_f(double v, int w, int d) {
    char b[400];
    snprintf(b, sizeof b, "%*.*f", w, d, v);
    std::string s(b);
    if ((int)s.size() > w) s.assign((size_t)w, '*');
    return s;
}

// Ew.d (lead = 0: 0.ddd form) or ESw.d (lead = 1: d.ddd form); exponents beyond two digits drop the letter ("+100"), as the
// standard prescribes for an exponent field without a width
inline std::string fmt_exp(double v, int w, int d, int lead) {
    std::string s;
    if (v != v) s = "NaN";
    else if (std::isinf(v)) s = v > 0 ? "Infinity" : "-Infinity";
    else {
        const int nsig = d + lead;                               // significant digits
        char b[64];
        int ex = 0;
        std::string digits((size_t)nsig, '0');
        if (v != 0.0) {
            snprintf(b, sizeof b, "%.*e", nsig - 1, std::fabs(v));   // D.DDDDe+XX, correctly rounded
            const char *pe = strchr(b, 'e');
            ex = atoi(pe + 1) + (lead ? 0 : 1);
            digits = std::string(1, b[0]) + std::string(b + 2, (size_t)(pe - b - 2));
        }
        char e[16];
        const int ax = ex < 0 ? -ex : ex;
        if (ax <= 99) snprintf(e, sizeof e, "E%c%02d", ex < 0 ? '-' : '+', ax);
        else snprintf(e, sizeof e, "%c%03d", ex < 0 ? '-' : '+', ax);
        s = (std::signbit(v) && v != 0.0) ? "-" : "";
        s += lead ? digits.substr(0, 1) + "." + digits.substr(1) : "0." + digits;
        s += e;
    }
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)w - s.size(), ' ') + s;
}
inline std::string fmt_e(double v) { return fmt_exp(v, 25, 16, 0); }     // '(e)'
inline std::string fmt_es(double v) { return fmt_exp(v, 25, 16, 1); }    // '(es)'
inline std::string fmt_fd(double v) { return fmt_f(v, 25, 16); }         // '(f)'

// One number of a formatted or list-directed Fortran record: D exponents, and exponents without a letter ("0.1+100")
inline bool parse_fortran_real(const char *p, const char *end, double &out) {
    char b[64];
    size_t n = 0;
    bool seen_digit = false, seen_exp = false;
    for (; p < end && n + 2 < sizeof b; ++p) {
        char ch = *p;
        if (seen_digit && (ch == 'D' || ch == 'd' || ch == 'e')) ch = 'E';
        if (seen_digit && ch == 'E') seen_exp = true;
        if ((ch == '+' || ch == '-') && seen_digit && !seen_exp && b[n - 1] != 'E') { b[n++] = 'E'; seen_exp = true; }
        if ((ch >= '0' && ch <= '9') || ch == '.') seen_digit = true;
        b[n++] = ch;
    }
    b[n] = 0;
    if (n == 0) return false;                     // "Infinity" / "NaN" (no digit) go to strtod as they are
    char *q = nullptr;
    out = strtod(b, &q);
    return q && *q == 0;
}

// The numbers of one line (blank- or comma-separated).  false: a token is not a number.
inline bool parse_fortran_line(const std::string &line, std::vector<double> &out) {
    out.clear();
    const char *p = line.c_str(), *end = p + line.size();
    while (p < end) {
        while (p < end && (*p == ' ' || *p == '\t' || *p == ',' || *p == '\r')) ++p;
        if (p >= end) break;
        const char *q = p;
        while (q < end && *q != ' ' && *q != '\t' && *q != ',' && *q != '\r') ++q;
        double v;
        if (!parse_fortran_real(p, q, v)) return false;
        out.push_back(v);
        p = q;
    }
    return true;
}

}  // namespace trk3
