#include "Json.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace Json {

static const Value kNull;

const Value& Value::operator[](const std::string& key) const {
    if (mType != objectValue)
        return kNull;
    auto it = mObj.find(key);
    return it == mObj.end() ? kNull : it->second;
}

Value& Value::operator[](const std::string& key) {
    if (mType != objectValue) {
        *this = Value(objectValue);
    }
    return mObj[key];
}

std::vector<std::string> Value::getMemberNames() const {
    std::vector<std::string> out;
    if (mType == objectValue)
        for (auto& kv : mObj)
            out.push_back(kv.first);
    return out;
}

Value& Value::append(const Value& v) {
    if (mType != arrayValue)
        *this = Value(arrayValue);
    mArr.push_back(v);
    return mArr.back();
}

static void write_string(const std::string& s, std::string& out) {
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"':
                out += "\\\"";
                break;
            case '\\':
                out += "\\\\";
                break;
            case '\n':
                out += "\\n";
                break;
            case '\r':
                out += "\\r";
                break;
            case '\t':
                out += "\\t";
                break;
            case '\b':
                out += "\\b";
                break;
            case '\f':
                out += "\\f";
                break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof buf, "\\u%04x", c);
                    out += buf;
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
}

void Value::write(std::string& out) const {
    char buf[64];
    switch (mType) {
        case nullValue:
            out += "null";
            break;
        case booleanValue:
            out += mInt ? "true" : "false";
            break;
        case intValue:
            snprintf(buf, sizeof buf, "%lld", (long long)mInt);
            out += buf;
            break;
        case uintValue:
            snprintf(buf, sizeof buf, "%llu", (unsigned long long)mInt);
            out += buf;
            break;
        case realValue:
            snprintf(buf, sizeof buf, "%.17g", mReal);
            out += buf;
            break;
        case stringValue:
            write_string(mStr, out);
            break;
        case arrayValue: {
            out.push_back('[');
            bool first = true;
            for (auto& v : mArr) {
                if (!first)
                    out.push_back(',');
                first = false;
                v.write(out);
            }
            out.push_back(']');
            break;
        }
        case objectValue: {
            out.push_back('{');
            bool first = true;
            for (auto& kv : mObj) {
                if (!first)
                    out.push_back(',');
                first = false;
                write_string(kv.first, out);
                out.push_back(':');
                kv.second.write(out);
            }
            out.push_back('}');
            break;
        }
    }
}

std::string Value::toString() const {
    std::string out;
    write(out);
    return out;
}

namespace {
struct P {
    const char* p;
    const char* e;
    std::string err;
    void ws() {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r'))
            ++p;
    }
    bool fail(const char* m) {
        if (err.empty())
            err = m;
        return false;
    }
    static void utf8(uint32_t cp, std::string& out) {
        if (cp < 0x80) {
            out.push_back((char)cp);
        } else if (cp < 0x800) {
            out.push_back((char)(0xC0 | (cp >> 6)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18)));
            out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    bool hex4(uint32_t& v) {
        if (e - p < 4)
            return fail("bad \\u escape");
        v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9')
                v |= c - '0';
            else if (c >= 'a' && c <= 'f')
                v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F')
                v |= c - 'A' + 10;
            else
                return fail("bad \\u escape");
        }
        return true;
    }
    bool str(std::string& out) {
        ++p; // opening quote
        while (p < e && *p != '"') {
            if (*p == '\\') {
                ++p;
                if (p >= e)
                    return fail("bad escape");
                char c = *p++;
                switch (c) {
                    case 'n':
                        out.push_back('\n');
                        break;
                    case 't':
                        out.push_back('\t');
                        break;
                    case 'r':
                        out.push_back('\r');
                        break;
                    case 'b':
                        out.push_back('\b');
                        break;
                    case 'f':
                        out.push_back('\f');
                        break;
                    case 'u': {
                        uint32_t cp;
                        if (!hex4(cp))
                            return false;
                        if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            uint32_t lo;
                            if (!hex4(lo))
                                return false;
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        utf8(cp, out);
                        break;
                    }
                    default:
                        out.push_back(c);
                }
            } else {
                out.push_back(*p++); // raw control characters (e.g. embedded newlines) are accepted
            }
        }
        if (p >= e)
            return fail("unterminated string");
        ++p;
        return true;
    }
    bool value(Value& out) {
        ws();
        if (p >= e)
            return fail("unexpected end");
        char c = *p;
        if (c == '{') {
            ++p;
            out = Value(objectValue);
            ws();
            if (p < e && *p == '}') {
                ++p;
                return true;
            }
            for (;;) {
                ws();
                if (p >= e || *p != '"')
                    return fail("expected key");
                std::string k;
                if (!str(k))
                    return false;
                ws();
                if (p >= e || *p != ':')
                    return fail("expected ':'");
                ++p;
                Value v;
                if (!value(v))
                    return false;
                out[k] = v;
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == '}') {
                    ++p;
                    return true;
                }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p;
            out = Value(arrayValue);
            ws();
            if (p < e && *p == ']') {
                ++p;
                return true;
            }
            for (;;) {
                Value v;
                if (!value(v))
                    return false;
                out.append(v);
                ws();
                if (p < e && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < e && *p == ']') {
                    ++p;
                    return true;
                }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            std::string s;
            if (!str(s))
                return false;
            out = Value(s);
            return true;
        }
        if (e - p >= 4 && !strncmp(p, "true", 4)) {
            p += 4;
            out = Value(true);
            return true;
        }
        if (e - p >= 5 && !strncmp(p, "false", 5)) {
            p += 5;
            out = Value(false);
            return true;
        }
        if (e - p >= 4 && !strncmp(p, "null", 4)) {
            p += 4;
            out = Value();
            return true;
        }
        // number
        const char* s = p;
        bool real = false;
        if (p < e && (*p == '-' || *p == '+'))
            ++p;
        while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
            if (*p == '.' || *p == 'e' || *p == 'E')
                real = true;
            ++p;
        }
        if (p == s)
            return fail("unexpected character");
        std::string num(s, p - s);
        if (real)
            out = Value(strtod(num.c_str(), nullptr));
        else if (num[0] == '-')
            out = Value((int64_t)strtoll(num.c_str(), nullptr, 10));
        else
            out = Value((uint64_t)strtoull(num.c_str(), nullptr, 10));
        return true;
    }
};
} // namespace

bool Value::parse(const char* begin, const char* end, Value& out, std::string& err) {
    P ps{begin, end, {}};
    if (!ps.value(out)) {
        err = ps.err;
        return false;
    }
    ps.ws();
    if (ps.p != ps.e) {
        err = "trailing characters";
        return false;
    }
    return true;
}

} // namespace Json
