// A YAML reader for the subset the plugin configuration document uses
// (/root/reference/values.yaml:9-18): block mappings, block sequences (also at the parent key's own
// indentation), plain / single- / double-quoted scalars, comments, `key:` with a nested block, empty
// flow collections `[]` `{}`, and `|`/`|-`/`>` block scalars.  Scalars resolve like PyYAML's safe_load
// (the Python twin config.py uses it): null/~/empty -> null; true/false/yes/no/on/off in the three
// casings -> bool; decimal, 0x, 0o-less octal and sign -> int; everything else a string.
// Anchors, tags, multi-document streams and non-empty flow collections are rejected, not guessed at.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace yaml {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

struct Node {
    enum Kind { Null, Bool, Int, Str, Map, Seq } kind = Null;
    bool b = false;
    int64_t i = 0;
    std::string s;
    std::vector<std::pair<std::string, Node>> map;
    std::vector<Node> seq;

    const Node* get(const std::string& key) const {
        if (kind != Map) return nullptr;
        for (const auto& kv : map) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_null() const { return kind == Null; }
};

namespace detail {

struct Line { int indent; std::string text; int no; };

inline std::string rtrim(std::string s) { while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back(); return s; }

// strip a trailing comment that is outside quotes (a '#' preceded by whitespace or at column 0)
inline std::string strip_comment(const std::string& s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) {
            if (q == '"' && c == '\\') { ++i; continue; }
            if (c == q) { if (q == '\'' && i + 1 < s.size() && s[i + 1] == '\'') { ++i; continue; } q = 0; }
        } else if (c == '"' || c == '\'') {
            if (i == 0 || s[i - 1] == ' ' || s[i - 1] == ':' || s[i - 1] == '-') q = c;
        } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
            return rtrim(s.substr(0, i));
        }
    }
    return rtrim(s);
}

inline Node scalar(const std::string& raw, int line_no) {
    Node n;
    const std::string t = raw;
    if (t.empty() || t == "~" || t == "null" || t == "Null" || t == "NULL") return n;
    if (t[0] == '"' || t[0] == '\'') {
        const char q = t[0];
        if (t.size() < 2 || t.back() != q) throw Error("line " + std::to_string(line_no) + ": unterminated quoted scalar");
        n.kind = Node::Str;
        for (size_t i = 1; i + 1 < t.size(); ++i) {
            if (q == '\'' && t[i] == '\'' && i + 2 < t.size() && t[i + 1] == '\'') { n.s.push_back('\''); ++i; }
            else if (q == '"' && t[i] == '\\' && i + 2 < t.size()) {
                const char e = t[++i];
                n.s.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e == '0' ? '\0' : e);
            } else n.s.push_back(t[i]);
        }
        return n;
    }
    if (t[0] == '&' || t[0] == '*' || t[0] == '!') throw Error("line " + std::to_string(line_no) + ": anchors, aliases and tags are not supported");
    if (t == "[]") { n.kind = Node::Seq; return n; }
    if (t == "{}") { n.kind = Node::Map; return n; }
    if (t[0] == '[' || t[0] == '{') throw Error("line " + std::to_string(line_no) + ": non-empty flow collections are not supported");
    static const char* yes[] = {"true", "True", "TRUE", "yes", "Yes", "YES", "on", "On", "ON"};
    static const char* no[] = {"false", "False", "FALSE", "no", "No", "NO", "off", "Off", "OFF"};
    for (const char* y : yes) if (t == y) { n.kind = Node::Bool; n.b = true; return n; }
    for (const char* x : no) if (t == x) { n.kind = Node::Bool; n.b = false; return n; }
    // integers: [-+]? (0 | [1-9][0-9_]* | 0x[0-9a-fA-F_]+ | 0[0-7_]+)
    {
        size_t i = 0;
        bool neg = false;
        if (t[i] == '-' || t[i] == '+') { neg = t[i] == '-'; ++i; }
        if (i < t.size()) {
            int base = 10;
            size_t j = i;
            if (t.size() - j > 2 && t[j] == '0' && (t[j + 1] == 'x')) { base = 16; j += 2; }
            else if (t.size() - j > 1 && t[j] == '0') { base = 8; j += 1; }
            bool ok = j < t.size(), any = false;
            uint64_t v = 0;
            for (size_t k = j; k < t.size() && ok; ++k) {
                const char c = t[k];
                int d;
                if (c == '_') { if (!any) ok = false; continue; }
                if (c >= '0' && c <= '9') d = c - '0';
                else if (base == 16 && c >= 'a' && c <= 'f') d = c - 'a' + 10;
                else if (base == 16 && c >= 'A' && c <= 'F') d = c - 'A' + 10;
                else { ok = false; break; }
                if (d >= base) { ok = false; break; }
                v = v * (uint64_t)base + (uint64_t)d;
                any = true;
            }
            if (t.size() - i == 1 && t[i] == '0') { ok = true; any = true; v = 0; }
            if (ok && any) { n.kind = Node::Int; n.i = neg ? -(int64_t)v : (int64_t)v; return n; }
        }
    }
    n.kind = Node::Str;
    n.s = t;
    return n;
}

// position of the ": " (or trailing ':') that splits `key: value`, outside quotes; npos if none
inline size_t key_split(const std::string& s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (q) { if (c == q) q = 0; continue; }
        if ((c == '"' || c == '\'') && i == 0) { q = c; continue; }
        if (c == ':' && (i + 1 == s.size() || s[i + 1] == ' ')) return i;
    }
    return std::string::npos;
}

class Parser {
public:
    explicit Parser(const std::string& text) {
        size_t pos = 0;
        int no = 0;
        bool started = false;
        while (pos <= text.size()) {
            size_t e = text.find('\n', pos);
            if (e == std::string::npos) e = text.size();
            std::string raw = text.substr(pos, e - pos);
            pos = e + 1;
            ++no;
            if (raw.find('\t') != std::string::npos && raw.find_first_not_of(" \t") != std::string::npos &&
                raw.find('\t') < raw.find_first_not_of(" \t"))
                throw Error("line " + std::to_string(no) + ": tabs cannot indent YAML");
            int ind = 0;
            while ((size_t)ind < raw.size() && raw[ind] == ' ') ++ind;
            std::string body = rtrim(raw.substr(ind));
            if (body == "---") { if (started) throw Error("multi-document streams are not supported"); continue; }
            if (body == "...") break;
            lines_.push_back({ind, body, no});                // comments are stripped lazily: block scalars keep '#'
            if (!body.empty() && body[0] != '#') started = true;
            if (e == text.size()) break;
        }
    }
    Node parse() {
        skip_blank();
        if (i_ >= lines_.size()) return Node();
        Node n = block(lines_[i_].indent);
        skip_blank();
        if (i_ < lines_.size()) throw Error("line " + std::to_string(lines_[i_].no) + ": unexpected content (bad indentation?)");
        return n;
    }

private:
    void skip_blank() {
        while (i_ < lines_.size()) {
            const std::string t = strip_comment(lines_[i_].text);
            if (!t.empty()) break;
            ++i_;
        }
    }
    static bool is_item(const std::string& t) { return t == "-" || (t.size() >= 2 && t[0] == '-' && t[1] == ' '); }

    Node block(int indent) {
        skip_blank();
        const std::string t = strip_comment(lines_[i_].text);
        return is_item(t) ? sequence(indent) : mapping(indent);
    }
    // value that follows "key:" (or "- ") on the same line, or the nested block below it
    Node value_after(const std::string& rest_in, int parent_indent, int line_no, bool parent_is_map) {
        const std::string rest = rest_in;
        if (!rest.empty() && (rest[0] == '|' || rest[0] == '>')) return block_scalar(rest, parent_indent);
        if (!rest.empty()) return scalar(rest, line_no);
        skip_blank();
        if (i_ >= lines_.size()) return Node();
        const int ind = lines_[i_].indent;
        const std::string t = strip_comment(lines_[i_].text);
        if (ind > parent_indent) return block(ind);
        if (parent_is_map && ind == parent_indent && is_item(t)) return sequence(ind);     // "key:\n- a\n- b"
        return Node();
    }
    Node mapping(int indent) {
        Node n;
        n.kind = Node::Map;
        for (;;) {
            skip_blank();
            if (i_ >= lines_.size() || lines_[i_].indent != indent) {
                if (i_ < lines_.size() && lines_[i_].indent > indent) throw Error("line " + std::to_string(lines_[i_].no) + ": bad indentation of a mapping entry");
                break;
            }
            const Line ln = lines_[i_];
            const std::string t = strip_comment(ln.text);
            if (is_item(t)) break;
            const size_t c = key_split(t);
            if (c == std::string::npos) throw Error("line " + std::to_string(ln.no) + ": expected 'key: value'");
            Node k = scalar(rtrim(t.substr(0, c)), ln.no);
            std::string key = k.kind == Node::Str ? k.s : rtrim(t.substr(0, c));
            for (const auto& kv : n.map) if (kv.first == key) throw Error("line " + std::to_string(ln.no) + ": duplicate key '" + key + "'");
            std::string rest = c + 1 < t.size() ? t.substr(c + 1) : "";
            rest.erase(0, rest.find_first_not_of(' ') == std::string::npos ? rest.size() : rest.find_first_not_of(' '));
            ++i_;
            n.map.emplace_back(key, value_after(rest, indent, ln.no, true));
        }
        return n;
    }
    Node sequence(int indent) {
        Node n;
        n.kind = Node::Seq;
        for (;;) {
            skip_blank();
            if (i_ >= lines_.size() || lines_[i_].indent != indent) break;
            const Line ln = lines_[i_];
            const std::string t = strip_comment(ln.text);
            if (!is_item(t)) break;
            std::string rest = t.size() > 2 ? t.substr(2) : "";
            const size_t lead = rest.find_first_not_of(' ');
            const int inner = indent + 2 + (lead == std::string::npos ? 0 : (int)lead);
            rest = lead == std::string::npos ? "" : rest.substr(lead);
            if (!rest.empty() && !is_item(rest) && key_split(rest) != std::string::npos && rest[0] != '"' && rest[0] != '\'') {
                // "- key: value" opens a mapping whose entries continue at column `inner`
                lines_[i_].indent = inner;
                lines_[i_].text = rest;
                n.seq.push_back(mapping(inner));
            } else if (is_item(rest)) {
                lines_[i_].indent = inner;
                lines_[i_].text = rest;
                n.seq.push_back(sequence(inner));
            } else {
                ++i_;
                n.seq.push_back(value_after(rest, indent, ln.no, false));
            }
        }
        return n;
    }
    // "|", "|-", "|+", ">" …: lines more indented than the parent, verbatim (comments included)
    Node block_scalar(const std::string& header, int parent_indent) {
        const bool folded = header[0] == '>';
        char chomp = 0;
        for (size_t k = 1; k < header.size(); ++k) if (header[k] == '-' || header[k] == '+') chomp = header[k];
        int ind = -1;
        std::vector<std::string> out;
        while (i_ < lines_.size()) {
            const Line& ln = lines_[i_];
            if (ln.text.empty()) { out.push_back(""); ++i_; continue; }
            if (ln.indent <= parent_indent) break;
            if (ind < 0) ind = ln.indent;
            if (ln.indent < ind) break;
            out.push_back(std::string((size_t)(ln.indent - ind), ' ') + ln.text);
            ++i_;
        }
        size_t trailing = 0;
        while (!out.empty() && out.back().empty()) { out.pop_back(); ++trailing; }
        Node n;
        n.kind = Node::Str;
        for (size_t k = 0; k < out.size(); ++k) {
            n.s += out[k];
            if (k + 1 < out.size()) n.s += folded && !out[k].empty() && !out[k + 1].empty() ? " " : "\n";
        }
        if (!out.empty() && chomp != '-') n.s += "\n";
        if (chomp == '+') n.s += std::string(trailing, '\n');
        return n;
    }

    std::vector<Line> lines_;
    size_t i_ = 0;
};

}  // namespace detail

inline Node parse(const std::string& text) { return detail::Parser(text).parse(); }

}  // namespace yaml
