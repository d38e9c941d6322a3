// json_min.h — minimal JSON reader (RFC 8259 subset: no surrogate pairs beyond pass-through) shared by the glTF loader and the RTXPT material-file
// reader.  Errors are thrown as LoadError and turned into C ABI status codes at the boundary.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace rtxpt_host {

struct LoadError { std::string msg; };
[[noreturn]] inline void failf(const char* fmt, ...)
{
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    throw LoadError{ buf };
}

struct JValue
{
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    double num = 0; bool b = false; std::string str;
    std::vector<JValue> arr; std::vector<std::pair<std::string, JValue>> obj;
    const JValue* find(const char* key) const { if (type != Object) return nullptr; for (auto& kv : obj) if (kv.first == key) return &kv.second; return nullptr; }
    const JValue& at(const char* key) const { const JValue* v = find(key); if (!v) failf("glTF: missing property '%s'", key); return *v; }
    double number(const char* key, double def) const { const JValue* v = find(key); return (v && v->type == Number) ? v->num : def; }
    int integer(const char* key, int def) const { const JValue* v = find(key); return (v && v->type == Number) ? int(v->num) : def; }
    std::string string(const char* key, const char* def = "") const { const JValue* v = find(key); return (v && v->type == String) ? v->str : std::string(def); }
    size_t size() const { return type == Array ? arr.size() : 0; }
};
struct JParser
{
    const char* p; const char* end; int depth = 0;
    static constexpr int kMaxDepth = 256;            // nesting limit: a hostile file must not recurse the parser off the stack
    struct Nest { int& d; explicit Nest(int& dd) : d(dd) { if (++d > kMaxDepth) failf("JSON: nesting deeper than %d levels", kMaxDepth); } ~Nest() { --d; } };
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    JValue parse() { ws(); JValue v = value(); ws(); return v; }
    JValue value()
    {
        if (p >= end) failf("JSON: unexpected end");
        Nest nest(depth);
        JValue v;
        switch (*p)
        {
        case '{':
            v.type = JValue::Object; p++; ws();
            if (p < end && *p == '}') { p++; return v; }
            while (true)
            {
                ws(); if (p >= end || *p != '"') failf("JSON: expected string key");
                std::string k = str(); ws();
                if (p >= end || *p != ':') failf("JSON: expected ':'");
                p++; ws(); v.obj.emplace_back(std::move(k), value()); ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                failf("JSON: expected ',' or '}'");
            }
            return v;
        case '[':
            v.type = JValue::Array; p++; ws();
            if (p < end && *p == ']') { p++; return v; }
            while (true)
            {
                ws(); v.arr.push_back(value()); ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                failf("JSON: expected ',' or ']'");
            }
            return v;
        case '"': v.type = JValue::String; v.str = str(); return v;
        case 't': if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; v.type = JValue::Bool; v.b = true; return v; } break;
        case 'f': if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; v.type = JValue::Bool; v.b = false; return v; } break;
        case 'n': if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return v; } break;
        default:
        {
            // the number's characters are copied out first: the input buffer need not be NUL-terminated
            const char* q = p; while (q < end && q - p < 63 && ((*q >= '0' && *q <= '9') || *q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E')) q++;
            char tmp[64]; memcpy(tmp, p, size_t(q - p)); tmp[q - p] = 0;
            char* e = nullptr; v.num = strtod(tmp, &e);
            if (e == tmp) break;
            p += e - tmp; v.type = JValue::Number; return v;
        }
        }
        failf("JSON: unexpected character '%c'", *p);
    }
    std::string str()
    {
        std::string s; p++;
        while (p < end && *p != '"')
        {
            if (*p == '\\' && p + 1 < end)
            {
                p++;
                switch (*p)
                {
                case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break; case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                case 'u':
                {
                    if (end - p < 5) failf("JSON: bad \\u escape");
                    unsigned cp = unsigned(strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16)); p += 4;
                    if (cp < 0x80) s += char(cp);
                    else if (cp < 0x800) { s += char(0xC0 | (cp >> 6)); s += char(0x80 | (cp & 0x3F)); }
                    else { s += char(0xE0 | (cp >> 12)); s += char(0x80 | ((cp >> 6) & 0x3F)); s += char(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: s += *p; break;
                }
                p++;
            }
            else s += *p++;
        }
        if (p >= end) failf("JSON: unterminated string");
        p++;
        return s;
    }
};


} // namespace rtxpt_host
