// csrc/json.h -- minimal JSON reader for the octvr camera configs (schema: apps/octvr/dump.cpp:71-96,
// modules/octvr/src/camera.cpp:49-135 and the per-model constructors under src/cameras/).
#pragma once
#include <charconv>
#include "common.h"
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>

namespace ob {

struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    bool has(const std::string& k) const { return find(k) != nullptr; }
    const Json* find(const std::string& k) const
    {
        if (kind != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Json& at(const std::string& k) const
    {
        const Json* j = find(k);
        if (!j) fail(OCTVR_ERR_FORMAT, "config: missing key \"" + k + "\"");
        return *j;
    }
    const Json& at(size_t i) const
    {
        if (kind != Arr || i >= arr.size()) fail(OCTVR_ERR_FORMAT, "config: array index out of range");
        return arr[i];
    }
    double number() const
    {
        if (kind != Num) fail(OCTVR_ERR_FORMAT, "config: number expected");
        return num;
    }
    int integer() const { return (int)number(); }
    bool boolean() const
    {
        if (kind != Bool) fail(OCTVR_ERR_FORMAT, "config: bool expected");
        return b;
    }
    const std::string& string() const
    {
        if (kind != Str) fail(OCTVR_ERR_FORMAT, "config: string expected");
        return str;
    }
    size_t size() const { return kind == Arr ? arr.size() : kind == Obj ? obj.size() : 0; }
};

class JsonParser {
    const char* p; const char* e;
    void ws() { while (p < e && std::isspace((unsigned char)*p)) p++; }
    [[noreturn]] void bad(const char* m) { fail(OCTVR_ERR_FORMAT, std::string("config: JSON parse error: ") + m); }
    int depth = 0;
    struct Nest { int& d; explicit Nest(int& d_) : d(d_) { ++d; } ~Nest() { --d; } };
    Json value()
    {
        Nest nest(depth);
        if (depth > 64) bad("nesting deeper than 64 levels");          // recursion bound: a hostile config must not overflow the stack
        ws();
        if (p >= e) bad("unexpected end");
        Json j;
        if (*p == '{') {
            j.kind = Json::Obj; p++; ws();
            if (p < e && *p == '}') { p++; return j; }
            for (;;) {
                ws();
                Json k = value();
                if (k.kind != Json::Str) bad("object key must be a string");
                ws();
                if (p >= e || *p != ':') bad("':' expected");
                p++;
                j.obj.emplace_back(k.str, value());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; break; }
                bad("',' or '}' expected");
            }
        } else if (*p == '[') {
            j.kind = Json::Arr; p++; ws();
            if (p < e && *p == ']') { p++; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; break; }
                bad("',' or ']' expected");
            }
        } else if (*p == '"') {
            j.kind = Json::Str; p++;
            while (p < e && *p != '"') {
                if (*p == '\\' && p + 1 < e) {
                    p++;
                    switch (*p) { case 'n': j.str += '\n'; break; case 't': j.str += '\t'; break; case 'r': j.str += '\r'; break;
                                  case 'b': j.str += '\b'; break; case 'f': j.str += '\f'; break;
                                  case 'u': if (p + 4 < e) { j.str += '?'; p += 4; } break;
                                  default: j.str += *p; }
                    p++;
                } else j.str += *p++;
            }
            if (p >= e) bad("unterminated string");
            p++;
        } else if (!strncmp_(p, "true")) { j.kind = Json::Bool; j.b = true; p += 4; }
        else if (!strncmp_(p, "false")) { j.kind = Json::Bool; j.b = false; p += 5; }
        else if (!strncmp_(p, "null")) { j.kind = Json::Null; p += 4; }
        else {
            // JSON number grammar only (no inf / nan / hex), parsed independently of the process locale (a host application
            // may have called setlocale): std::from_chars is correctly rounded like rapidjson's full-precision parse
            const char* q = p;
            if (q < e && *q == '-') q++;
            if (q >= e || !std::isdigit((unsigned char)*q)) bad("value expected");
            while (q < e && std::isdigit((unsigned char)*q)) q++;
            if (q < e && *q == '.') { q++; if (q >= e || !std::isdigit((unsigned char)*q)) bad("digit expected after '.'"); while (q < e && std::isdigit((unsigned char)*q)) q++; }
            if (q < e && (*q == 'e' || *q == 'E')) {
                q++;
                if (q < e && (*q == '+' || *q == '-')) q++;
                if (q >= e || !std::isdigit((unsigned char)*q)) bad("digit expected in exponent");
                while (q < e && std::isdigit((unsigned char)*q)) q++;
            }
            j.kind = Json::Num;
            const auto res = std::from_chars(p, q, j.num);
            if (res.ec != std::errc() || res.ptr != q) bad("number out of range");
            p = q;
        }
        return j;
    }
    int strncmp_(const char* a, const char* lit)
    {
        size_t n = std::char_traits<char>::length(lit);
        if ((size_t)(e - a) < n) return 1;
        return std::char_traits<char>::compare(a, lit, n);
    }
public:
    static Json parse(const std::string& s)
    {
        JsonParser q;
        q.p = s.data(); q.e = s.data() + s.size();
        Json j = q.value();
        q.ws();
        if (q.p != q.e) q.bad("trailing characters");
        return j;
    }
};

}  // namespace ob
