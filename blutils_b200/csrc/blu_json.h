// blu_json.h -- minimal JSON pull parser + serde_json-compatible emit helpers (host only).
// Used for the `.blutils.json` taxonomy file (serde_json::from_str::<TaxonomiesMap>, reference
// core/src/use_cases/build_consensus_identities/mod.rs:254-265), the custom cutoff file
// (core/src/domain/dtos/taxon.rs:28-65) and the writer (core/src/use_cases/write_blutils_output.rs).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>

namespace blu {

struct JsonError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct JsonCursor {
    const char* p;
    const char* e;
    const char* base;
    JsonCursor(const char* b, size_t n) : p(b), e(b + n), base(b) {}

    [[noreturn]] void fail(const char* what) const {
        throw JsonError(std::string(what) + " at byte " + std::to_string(p - base));
    }
    void ws() {
        while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    char peek() {
        ws();
        if (p >= e) fail("unexpected end of JSON");
        return *p;
    }
    void expect(char c) {
        if (peek() != c) fail("unexpected character");
        p++;
    }
    bool consume(char c) {
        if (peek() == c) {
            p++;
            return true;
        }
        return false;
    }
    bool consume_null() {
        ws();
        if (e - p >= 4 && !memcmp(p, "null", 4)) {
            p += 4;
            return true;
        }
        return false;
    }
    static void put_utf8(std::string& o, uint32_t cp) {
        if (cp < 0x80)
            o.push_back((char)cp);
        else if (cp < 0x800) {
            o.push_back((char)(0xC0 | (cp >> 6)));
            o.push_back((char)(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            o.push_back((char)(0xE0 | (cp >> 12)));
            o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            o.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            o.push_back((char)(0xF0 | (cp >> 18)));
            o.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            o.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            o.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t hex4() {
        if (e - p < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9')
                v |= c - '0';
            else if (c >= 'a' && c <= 'f')
                v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F')
                v |= c - 'A' + 10;
            else
                fail("bad \\u escape");
        }
        return v;
    }
    // Parses a JSON string.  When `out` is null the content is skipped.
    void string(std::string* out) {
        expect('"');
        if (out) out->clear();
        while (true) {
            const char* q = p;
            while (q < e && *q != '"' && *q != '\\' && (unsigned char)*q >= 0x20) q++;
            if (out) out->append(p, q - p);
            p = q;
            if (p >= e) fail("unterminated string");
            if (*p == '"') {
                p++;
                return;
            }
            if ((unsigned char)*p < 0x20) fail("control character in string");
            p++;  // backslash
            if (p >= e) fail("unterminated escape");
            char c = *p++;
            uint32_t cp;
            switch (c) {
                case '"': cp = '"'; break;
                case '\\': cp = '\\'; break;
                case '/': cp = '/'; break;
                case 'b': cp = '\b'; break;
                case 'f': cp = '\f'; break;
                case 'n': cp = '\n'; break;
                case 'r': cp = '\r'; break;
                case 't': cp = '\t'; break;
                case 'u': {
                    cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00) {
                        if (e - p < 6 || p[0] != '\\' || p[1] != 'u') fail("lone surrogate");
                        p += 2;
                        uint32_t lo = hex4();
                        if (lo < 0xDC00 || lo > 0xDFFF) fail("bad surrogate pair");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    } else if (cp >= 0xDC00 && cp < 0xE000)
                        fail("lone surrogate");
                    break;
                }
                default: fail("bad escape");
            }
            if (out) put_utf8(*out, cp);
        }
    }
    // Non-negative integer literal that fits u64 (serde: u64 field)
    uint64_t u64() {
        ws();
        const char* s = p;
        if (p < e && *p == '-') fail("expected an unsigned integer");
        uint64_t v = 0;
        while (p < e && *p >= '0' && *p <= '9') {
            uint64_t d = (uint64_t)(*p - '0');
            if (v > (UINT64_MAX - d) / 10) fail("integer overflows u64");
            v = v * 10 + d;
            p++;
        }
        if (p == s) fail("expected an unsigned integer");
        if (p < e && (*p == '.' || *p == 'e' || *p == 'E')) fail("expected an unsigned integer");
        return v;
    }
    int64_t i64() {
        ws();
        bool neg = false;
        if (p < e && *p == '-') {
            neg = true;
            p++;
        }
        uint64_t v = u64();
        if (v > (uint64_t)INT64_MAX) fail("integer overflows i64");
        return neg ? -(int64_t)v : (int64_t)v;
    }
    void skip_value() {
        char c = peek();
        if (c == '"') {
            string(nullptr);
        } else if (c == '{') {
            p++;
            if (consume('}')) return;
            do {
                string(nullptr);
                expect(':');
                skip_value();
            } while (consume(','));
            expect('}');
        } else if (c == '[') {
            p++;
            if (consume(']')) return;
            do {
                skip_value();
            } while (consume(','));
            expect(']');
        } else {
            const char* s = p;
            while (p < e && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
            std::string_view t(s, p - s);
            if (t == "null" || t == "true" || t == "false") return;
            if (t.empty()) fail("expected a value");
            for (char ch : t)
                if (!((ch >= '0' && ch <= '9') || ch == '-' || ch == '+' || ch == '.' || ch == 'e' || ch == 'E')) fail("bad literal");
        }
    }
    void end() {
        ws();
        if (p != e) fail("trailing characters");
    }
};

// ---- emit (serde_json) ------------------------------------------------------------------------------------
inline void json_escape(std::string& o, std::string_view s) {
    o.push_back('"');
    size_t run = 0;
    const char* d = s.data();
    for (size_t i = 0; i < s.size(); i++) {
        unsigned char c = (unsigned char)d[i];
        if (c >= 0x20 && c != '"' && c != '\\') continue;
        o.append(d + run, i - run);
        run = i + 1;
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default: {
                static const char* hex = "0123456789abcdef";
                o += "\\u00";
                o.push_back(hex[c >> 4]);
                o.push_back(hex[c & 15]);
            }
        }
    }
    o.append(d + run, s.size() - run);
    o.push_back('"');
}

// serde_json f64: ryu shortest digits in ryu's "pretty" layout; non-finite -> null.
inline void json_f64(std::string& o, double v) {
    if (!std::isfinite(v)) {
        o += "null";
        return;
    }
    if (v == 0) {
        o += std::signbit(v) ? "-0.0" : "0.0";
        return;
    }
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    const char* ep = buf;
    while (ep < r.ptr && *ep != 'e') ep++;
    int ex = 0;
    std::from_chars(ep + (ep[1] == '+' ? 2 : 1), r.ptr, ex);
    char dg[24];
    int n = 0;
    for (const char* c = buf; c < ep; c++)
        if (*c != '.') dg[n++] = *c;
    while (n > 1 && dg[n - 1] == '0') n--;
    const int kk = ex + 1, k = kk - n;
    if (std::signbit(v)) o.push_back('-');
    if (0 <= k && kk <= 16) {
        o.append(dg, n);
        o.append((size_t)k, '0');
        o += ".0";
    } else if (0 < kk && kk <= 16) {
        o.append(dg, kk);
        o.push_back('.');
        o.append(dg + kk, n - kk);
    } else if (-5 < kk && kk <= 0) {
        o += "0.";
        o.append((size_t)(-kk), '0');
        o.append(dg, n);
    } else {
        o.push_back(dg[0]);
        if (n > 1) {
            o.push_back('.');
            o.append(dg + 1, n - 1);
        }
        o.push_back('e');
        o += std::to_string(kk - 1);
    }
}

// Rust `format!("{}", f64)`: shortest round-trip digits, positional notation, no exponent, no trailing ".0"
// (used by parse_consensus_as_tabular, reference core/src/use_cases/parse_consensus_as_tabular/mod.rs:133-134)
inline void rust_display_f64(std::string& o, double v) {
    if (std::isnan(v)) {
        o += "NaN";
        return;
    }
    if (std::isinf(v)) {
        o += v < 0 ? "-inf" : "inf";
        return;
    }
    if (v == 0) {
        o += std::signbit(v) ? "-0" : "0";
        return;
    }
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    const char* ep = buf;
    while (ep < r.ptr && *ep != 'e') ep++;
    int ex = 0;
    std::from_chars(ep + (ep[1] == '+' ? 2 : 1), r.ptr, ex);
    char dg[24];
    int n = 0;
    for (const char* c = buf; c < ep; c++)
        if (*c != '.') dg[n++] = *c;
    while (n > 1 && dg[n - 1] == '0') n--;
    const int kk = ex + 1;  // digits before the decimal point
    if (std::signbit(v)) o.push_back('-');
    if (kk <= 0) {
        o += "0.";
        o.append((size_t)(-kk), '0');
        o.append(dg, n);
    } else if (kk >= n) {
        o.append(dg, n);
        o.append((size_t)(kk - n), '0');
    } else {
        o.append(dg, kk);
        o.push_back('.');
        o.append(dg + kk, n - kk);
    }
}

}  // namespace blu
