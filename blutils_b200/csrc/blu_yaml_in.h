// blu_yaml_in.h -- reader for the YAML that `blu blastn build-consensus --out-format yaml` writes (serde_yaml 0.9 block style),
// used by build-tabular (reference: FileOrStdin::yaml_content -> serde_yaml::from_str::<BlutilsOutput>,
// core/src/domain/dtos/file_or_stdin.rs:113-130).  Host only.
//
// The reference accepts any YAML 1.2 document through unsafe-libyaml; restating a complete YAML parser is out of
// proportion for a result file whose shape is fixed, so this reader covers the block-style subset such files are written in --
// and everything a person is likely to do to one by hand -- and refuses the rest LOUDLY (YamlUnsupported ->
// BLU_ERR_UNSUPPORTED), never guessing:
//   supported    block mappings and block sequences by indentation (a sequence may sit at its key's own indentation, as
//                serde_yaml writes it), `- key: value` items, plain / 'single' / "double" quoted scalars on one line with
//                every YAML escape, `[]` and `{}`, comments, blank lines, one leading `---`, a trailing `...`
//   refused      flow collections with content, block scalars (`|`, `>`), multi-line scalars, anchors / aliases, tags,
//                `? ` complex keys, a second document, tabs in indentation
// Typing follows serde_yaml's typed deserialisation: a String field takes any scalar verbatim; Option is None for a PLAIN
// `null` / `Null` / `NULL` / `~` / empty; bool is a plain true|True|TRUE|false|False|FALSE; numbers are plain scalars of the
// YAML 1.2 core schema (ints: decimal, 0x, 0o, 0b, optional sign, no leading zeros; floats add . e .inf .nan).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace blu {

struct YamlError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct YamlUnsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct YNode {
    enum Kind { Scalar, Map, Seq } kind = Scalar;
    std::string text;    // Scalar: its value (escapes resolved)
    bool plain = true;   // Scalar: not quoted (subject to null / bool / number typing)
    int line = 0;
    std::vector<std::pair<std::string, std::unique_ptr<YNode>>> map;
    std::vector<std::unique_ptr<YNode>> seq;
};

class YamlReader {
    struct Line {
        int indent;
        std::string_view body;  // content from the first non-space character, trailing spaces removed
        int no;
    };
    std::vector<Line> lines_;
    size_t at_ = 0;
    // streaming: the items of the block sequence under this root key are handed to the callback one by one and not kept
    // (a result file holds millions of them; a tree of all of them would be several times the file)
    std::string stream_key_;
    std::function<void(const YNode&)> on_item_;
    bool streaming_next_seq_ = false;

    [[noreturn]] static void fail(int line, const std::string& m) { throw YamlError(m + " at line " + std::to_string(line)); }
    [[noreturn]] static void refuse(int line, const std::string& m) { throw YamlUnsupported("YAML input: " + m + " at line " + std::to_string(line) + " is not supported by this reader"); }

    static void append_utf8(std::string& o, uint32_t cp) {
        if (cp < 0x80)
            o.push_back((char)cp);
        else if (cp < 0x800)
            o.push_back((char)(0xC0 | (cp >> 6))), o.push_back((char)(0x80 | (cp & 0x3F)));
        else if (cp < 0x10000)
            o.push_back((char)(0xE0 | (cp >> 12))), o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))), o.push_back((char)(0x80 | (cp & 0x3F)));
        else
            o.push_back((char)(0xF0 | (cp >> 18))), o.push_back((char)(0x80 | ((cp >> 12) & 0x3F))), o.push_back((char)(0x80 | ((cp >> 6) & 0x3F))),
                o.push_back((char)(0x80 | (cp & 0x3F)));
    }

    // a quoted scalar starting at s[0] (' or "): value -> out, returns the number of bytes consumed (including both quotes)
    static size_t quoted(std::string_view s, int line, std::string& out) {
        const char q = s[0];
        out.clear();
        size_t i = 1;
        for (; i < s.size(); i++) {
            const char c = s[i];
            if (q == '\'') {
                if (c == '\'') {
                    if (i + 1 < s.size() && s[i + 1] == '\'') {
                        out.push_back('\'');
                        i++;
                        continue;
                    }
                    return i + 1;
                }
                out.push_back(c);
                continue;
            }
            if (c == '"') return i + 1;
            if (c != '\\') {
                out.push_back(c);
                continue;
            }
            if (++i >= s.size()) refuse(line, "a double-quoted scalar that continues on the next line");
            const char e = s[i];
            auto hex = [&](int n) {
                if (i + (size_t)n >= s.size()) fail(line, "truncated escape in a double-quoted scalar");
                uint32_t v = 0;
                for (int k = 1; k <= n; k++) {
                    const char h = s[i + (size_t)k];
                    const int d = h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : -1;
                    if (d < 0) fail(line, "did not find expected hexdecimal number");
                    v = v * 16 + (uint32_t)d;
                }
                i += (size_t)n;
                if ((v >= 0xD800 && v <= 0xDFFF) || v > 0x10FFFF) fail(line, "found invalid Unicode character escape code");
                append_utf8(out, v);
            };
            switch (e) {
                case '0': out.push_back('\0'); break;
                case 'a': out.push_back('\a'); break;
                case 'b': out.push_back('\b'); break;
                case 't': case '\t': out.push_back('\t'); break;
                case 'n': out.push_back('\n'); break;
                case 'v': out.push_back('\v'); break;
                case 'f': out.push_back('\f'); break;
                case 'r': out.push_back('\r'); break;
                case 'e': out.push_back('\x1b'); break;
                case ' ': out.push_back(' '); break;
                case '"': out.push_back('"'); break;
                case '/': out.push_back('/'); break;
                case '\\': out.push_back('\\'); break;
                case 'N': append_utf8(out, 0x85); break;
                case '_': append_utf8(out, 0xA0); break;
                case 'L': append_utf8(out, 0x2028); break;
                case 'P': append_utf8(out, 0x2029); break;
                case 'x': hex(2); break;
                case 'u': hex(4); break;
                case 'U': hex(8); break;
                default: fail(line, "found unknown escape character");
            }
        }
        refuse(line, "a quoted scalar that continues on the next line");
    }

    // what follows a complete value on its line: nothing, or a comment
    static void only_comment(std::string_view rest, int line) {
        size_t i = 0;
        while (i < rest.size() && rest[i] == ' ') i++;
        if (i == rest.size()) return;
        if (rest[i] == '#' && i > 0) return;
        fail(line, "did not find expected key / end of the scalar");
    }

    static std::string_view strip_comment(std::string_view s) {  // plain text: a '#' behind a space starts a comment
        for (size_t i = 1; i < s.size(); i++)
            if (s[i] == '#' && s[i - 1] == ' ') {
                s = s.substr(0, i);
                break;
            }
        while (!s.empty() && s.back() == ' ') s.remove_suffix(1);
        return s;
    }

    // an inline value (the text behind "key: " or "- "), never empty
    std::unique_ptr<YNode> inline_value(std::string_view v, int line) {
        auto n = std::make_unique<YNode>();
        n->line = line;
        const char c = v[0];
        if (c == '"' || c == '\'') {
            const size_t used = quoted(v, line, n->text);
            only_comment(v.substr(used), line);
            n->plain = false;
            return n;
        }
        if (c == '[' || c == '{') {
            std::string_view t = strip_comment(v);
            std::string compact;
            for (char ch : t)
                if (ch != ' ') compact.push_back(ch);
            if (compact == "[]") {
                n->kind = YNode::Seq;
                return n;
            }
            if (compact == "{}") {
                n->kind = YNode::Map;
                return n;
            }
            refuse(line, "a flow collection with content");
        }
        if (c == '|' || c == '>') refuse(line, "a block scalar");
        if (c == '&' || c == '*') refuse(line, "an anchor / alias");
        if (c == '!') refuse(line, "a tag");
        if (c == '%' || c == '@' || c == '`') fail(line, "found character that cannot start any token");
        if (c == '#') {  // only a comment: the value is empty
            n->text.clear();
            return n;
        }
        std::string_view t = strip_comment(v);
        if (t.find(": ") != std::string_view::npos || (!t.empty() && t.back() == ':')) fail(line, "mapping values are not allowed in this context");
        n->text.assign(t);
        return n;
    }

    // splits "key: value" / "key:"; false when the line is not a mapping entry
    static bool split_key(std::string_view body, int line, std::string& key, std::string_view& value) {
        if (body.empty()) return false;
        size_t colon;
        if (body[0] == '"' || body[0] == '\'') {
            size_t used;
            try {
                used = quoted(body, line, key);
            } catch (const YamlUnsupported&) {
                return false;
            }
            size_t i = used;
            while (i < body.size() && body[i] == ' ') i++;
            if (i >= body.size() || body[i] != ':') return false;
            colon = i;
        } else {
            if (body[0] == '?' && (body.size() == 1 || body[1] == ' ')) refuse(line, "a complex key");
            colon = std::string_view::npos;
            for (size_t i = 0; i < body.size(); i++) {
                if (body[i] == '#' && i > 0 && body[i - 1] == ' ') break;
                if (body[i] == ':' && (i + 1 == body.size() || body[i + 1] == ' ')) {
                    colon = i;
                    break;
                }
            }
            if (colon == std::string_view::npos || colon == 0) return false;
            std::string_view k = body.substr(0, colon);
            while (!k.empty() && k.back() == ' ') k.remove_suffix(1);
            if (k[0] == '&' || k[0] == '*' || k[0] == '!') refuse(line, "an anchor / alias / tag");
            if (k[0] == '[' || k[0] == '{') refuse(line, "a flow collection as a key");
            key.assign(k);
        }
        value = body.substr(colon + 1);
        while (!value.empty() && value[0] == ' ') value.remove_prefix(1);
        return true;
    }

    static bool is_seq_item(std::string_view body) { return !body.empty() && body[0] == '-' && (body.size() == 1 || body[1] == ' '); }

    std::unique_ptr<YNode> null_node(int line) {
        auto n = std::make_unique<YNode>();
        n->line = line;
        return n;  // plain, empty: null
    }

    // the node that starts at the current line, which is indented by exactly `indent`
    std::unique_ptr<YNode> node(int indent, bool root = false) {
        const Line first = lines_[at_];
        const bool stream = streaming_next_seq_;  // (set by the root mapping for the value of the streamed key)
        streaming_next_seq_ = false;
        if (is_seq_item(first.body)) {
            auto n = std::make_unique<YNode>();
            n->kind = YNode::Seq;
            n->line = first.no;
            auto keep = [&](std::unique_ptr<YNode> item) {
                if (stream)
                    on_item_(*item);
                else
                    n->seq.push_back(std::move(item));
            };
            while (at_ < lines_.size() && lines_[at_].indent == indent && is_seq_item(lines_[at_].body)) {
                const Line cur = lines_[at_];
                std::string_view rest = cur.body.substr(1);
                size_t sp = 0;
                while (sp < rest.size() && rest[sp] == ' ') sp++;
                rest.remove_prefix(sp);
                if (rest.empty() || rest[0] == '#') {
                    at_++;
                    if (at_ < lines_.size() && lines_[at_].indent > indent)
                        keep(node(lines_[at_].indent));
                    else
                        keep(null_node(cur.no));
                    continue;
                }
                // the item's content behaves like a line of its own, indented to where it starts
                lines_[at_].indent = indent + 1 + (int)sp;
                lines_[at_].body = rest;
                keep(node(lines_[at_].indent));
            }
            if (at_ < lines_.size() && lines_[at_].indent > indent) fail(lines_[at_].no, "bad indentation of a sequence entry");
            return n;
        }
        std::string key;
        std::string_view value;
        if (split_key(first.body, first.no, key, value)) {
            auto n = std::make_unique<YNode>();
            n->kind = YNode::Map;
            n->line = first.no;
            while (at_ < lines_.size() && lines_[at_].indent == indent) {
                const Line cur = lines_[at_];
                if (is_seq_item(cur.body)) break;  // (a sequence at the indentation of the key it belongs to ends the PARENT's value)
                if (!split_key(cur.body, cur.no, key, value)) fail(cur.no, "could not find expected ':'");
                for (auto& kv : n->map)
                    if (kv.first == key) fail(cur.no, "duplicate entry with key \"" + key + "\"");
                at_++;
                std::unique_ptr<YNode> child;
                const bool streamed = root && on_item_ && key == stream_key_;
                if (!value.empty() && value[0] != '#')
                    child = inline_value(value, cur.no);
                else if (at_ < lines_.size() && lines_[at_].indent > indent) {
                    streaming_next_seq_ = streamed;
                    child = node(lines_[at_].indent);
                } else if (at_ < lines_.size() && lines_[at_].indent == indent && is_seq_item(lines_[at_].body)) {
                    streaming_next_seq_ = streamed;
                    child = node(indent);
                } else
                    child = null_node(cur.no);
                n->map.emplace_back(key, std::move(child));
            }
            if (at_ < lines_.size() && lines_[at_].indent > indent) fail(lines_[at_].no, "bad indentation of a mapping entry");
            return n;
        }
        // a scalar on a line of its own
        at_++;
        auto n = inline_value(first.body, first.no);
        if (n->kind == YNode::Scalar && n->plain && at_ < lines_.size() && lines_[at_].indent >= indent && indent > 0)
            refuse(lines_[at_].no, "a plain scalar that continues on the next line");
        return n;
    }

public:
    // The document's root node (a null scalar for an empty document).  With `stream_key` the items of the block sequence
    // under that key of the root mapping go to `on_item` as they are read and the returned tree holds an empty sequence there.
    std::unique_ptr<YNode> parse(std::string_view text, const std::string& stream_key = std::string(), std::function<void(const YNode&)> on_item = nullptr) {
        stream_key_ = stream_key;
        on_item_ = std::move(on_item);
        streaming_next_seq_ = false;
        lines_.clear();
        at_ = 0;
        int no = 0;
        bool started = false, ended = false;
        for (size_t pos = 0; pos <= text.size();) {
            size_t nl = text.find('\n', pos);
            if (nl == std::string_view::npos) nl = text.size();
            std::string_view ln = text.substr(pos, nl - pos);
            pos = nl + 1;
            no++;
            if (!ln.empty() && ln.back() == '\r') ln.remove_suffix(1);
            size_t ind = 0;
            while (ind < ln.size() && ln[ind] == ' ') ind++;
            std::string_view body = ln.substr(ind);
            while (!body.empty() && (body.back() == ' ' || body.back() == '\t')) body.remove_suffix(1);
            if (body.empty() || body[0] == '#') continue;
            if (body[0] == '\t') throw YamlError("found character that cannot start any token (a tab in the indentation) at line " + std::to_string(no));
            if (ended) refuse(no, "content behind the end of the document");
            if (ind == 0 && (body == "---" || body.substr(0, 4) == "--- ")) {
                if (started || !lines_.empty()) refuse(no, "a second document");
                started = true;
                body.remove_prefix(3);
                while (!body.empty() && body[0] == ' ') body.remove_prefix(1);
                if (body.empty() || body[0] == '#') continue;
                refuse(no, "content on the document start line");
            }
            if (ind == 0 && body == "...") {
                ended = true;
                continue;
            }
            if (ind == 0 && body[0] == '%') refuse(no, "a directive");
            lines_.push_back({(int)ind, body, no});
            if (pos > text.size()) break;
        }
        if (lines_.empty()) return null_node(1);
        auto root = node(lines_[0].indent, true);
        if (at_ < lines_.size()) fail(lines_[at_].no, "did not find expected <document end>");
        return root;
    }
};

// ---- serde_yaml's typed views of a scalar ----------------------------------------------------------------------------------
namespace yaml_typed {

[[noreturn]] inline void bad(const YNode& n, const std::string& m) { throw YamlError(m + " at line " + std::to_string(n.line)); }

inline bool is_null(const YNode& n) {
    return n.kind == YNode::Scalar && n.plain && (n.text.empty() || n.text == "~" || n.text == "null" || n.text == "Null" || n.text == "NULL");
}

inline const std::string& as_str(const YNode& n, const char* what) {
    if (n.kind != YNode::Scalar) bad(n, std::string("invalid type: expected a string for ") + what);
    return n.text;
}

inline bool as_bool(const YNode& n, const char* what) {
    if (n.kind == YNode::Scalar && n.plain) {
        const std::string& t = n.text;
        if (t == "true" || t == "True" || t == "TRUE") return true;
        if (t == "false" || t == "False" || t == "FALSE") return false;
    }
    bad(n, std::string("invalid type: expected a boolean for ") + what);
}

// YAML 1.2 core-schema integer (serde_yaml: optional sign, 0x / 0o / 0b, no leading zeros, no underscores)
inline bool parse_int(const std::string& t, int64_t& out) {
    size_t i = 0;
    bool neg = false;
    if (i < t.size() && (t[i] == '+' || t[i] == '-')) neg = t[i++] == '-';
    if (i >= t.size()) return false;
    int base = 10;
    if (t.size() - i > 2 && t[i] == '0' && (t[i + 1] == 'x' || t[i + 1] == 'o' || t[i + 1] == 'b')) {
        base = t[i + 1] == 'x' ? 16 : t[i + 1] == 'o' ? 8 : 2;
        i += 2;
    } else if (t.size() - i > 1 && t[i] == '0')
        return false;
    unsigned long long v = 0;
    for (; i < t.size(); i++) {
        const char c = t[i];
        const int d = c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99;
        if (d >= base) return false;
        if (v > (0x7FFFFFFFFFFFFFFFull + (neg ? 1ull : 0ull) - (unsigned)d) / (unsigned)base) return false;
        v = v * (unsigned)base + (unsigned)d;
    }
    out = neg ? (int64_t)(0ull - v) : (int64_t)v;
    return true;
}

inline int64_t as_i64(const YNode& n, const char* what) {
    int64_t v;
    if (n.kind == YNode::Scalar && n.plain && parse_int(n.text, v)) return v;
    bad(n, std::string("invalid type: expected an integer for ") + what);
}

inline double as_f64(const YNode& n, const char* what) {
    if (n.kind == YNode::Scalar && n.plain) {
        const std::string& t = n.text;
        int64_t iv;
        if (parse_int(t, iv)) return (double)iv;
        std::string_view u(t);
        bool neg = false;
        if (!u.empty() && (u[0] == '+' || u[0] == '-')) neg = u[0] == '-', u.remove_prefix(1);
        if (u == ".inf" || u == ".Inf" || u == ".INF") return neg ? -INFINITY : INFINITY;
        if (t == ".nan" || t == ".NaN" || t == ".NAN") return NAN;
        // [-+]? ( \. [0-9]+ | [0-9]+ ( \. [0-9]* )? ) ( [eE] [-+]? [0-9]+ )?
        size_t i = 0;
        size_t nd = 0;
        while (i < u.size() && u[i] >= '0' && u[i] <= '9') i++, nd++;
        if (i < u.size() && u[i] == '.') {
            i++;
            size_t nf = 0;
            while (i < u.size() && u[i] >= '0' && u[i] <= '9') i++, nf++;
            if (nd == 0 && nf == 0) nd = 0;
            else nd += nf;
        }
        if (nd > 0) {
            if (i < u.size() && (u[i] == 'e' || u[i] == 'E')) {
                i++;
                if (i < u.size() && (u[i] == '+' || u[i] == '-')) i++;
                size_t ne = 0;
                while (i < u.size() && u[i] >= '0' && u[i] <= '9') i++, ne++;
                if (ne == 0) i = u.size() + 1;
            }
            if (i == u.size()) {
                const double v = strtod(t.c_str(), nullptr);
                if (std::isfinite(v)) return v;  // (serde_yaml takes a literal that overflows to infinity for a string)
            }
        }
    }
    bad(n, std::string("invalid type: expected a float for ") + what);
}

}  // namespace yaml_typed
}  // namespace blu
