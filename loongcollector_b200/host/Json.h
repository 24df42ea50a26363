// Json.h -- a small JSON value in the shape of the subset of jsoncpp's Json::Value that the reference's
// processors and their unit tests use (Init(const Json::Value&), FromJsonString / ToJsonString).
// jsoncpp is not available in this image, so this is an independent minimal implementation.
#pragma once
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace Json {

enum ValueType { nullValue = 0, intValue, uintValue, realValue, stringValue, booleanValue, arrayValue, objectValue };

class Value {
public:
    Value() = default;
    Value(ValueType t) : mType(t) {}
    Value(bool b) : mType(booleanValue), mInt(b) {}
    Value(int v) : mType(intValue), mInt(v) {}
    Value(int64_t v) : mType(intValue), mInt(v) {}
    Value(uint64_t v) : mType(uintValue), mInt((int64_t)v) {}
    Value(double v) : mType(realValue), mReal(v) {}
    Value(const char* s) : mType(stringValue), mStr(s) {}
    Value(const std::string& s) : mType(stringValue), mStr(s) {}

    ValueType type() const { return mType; }
    bool isNull() const { return mType == nullValue; }
    bool isBool() const { return mType == booleanValue; }
    bool isInt() const { return mType == intValue || mType == uintValue; }
    bool isString() const { return mType == stringValue; }
    bool isArray() const { return mType == arrayValue; }
    bool isObject() const { return mType == objectValue; }

    bool asBool() const { return mInt != 0; }
    int asInt() const { return (int)mInt; }
    int64_t asInt64() const { return mInt; }
    uint64_t asUInt64() const { return (uint64_t)mInt; }
    const std::string& asString() const { return mStr; }

    bool isMember(const std::string& key) const { return mType == objectValue && mObj.count(key) != 0; }
    const Value& operator[](const std::string& key) const;
    Value& operator[](const std::string& key);
    const Value& operator[](const char* key) const { return (*this)[std::string(key)]; }
    Value& operator[](const char* key) { return (*this)[std::string(key)]; }
    std::vector<std::string> getMemberNames() const; // sorted, like jsoncpp
    size_t size() const { return mType == arrayValue ? mArr.size() : (mType == objectValue ? mObj.size() : 0); }
    const Value& operator[](size_t i) const { return mArr[i]; }
    Value& append(const Value& v);
    std::vector<Value>::const_iterator begin() const { return mArr.begin(); }
    std::vector<Value>::const_iterator end() const { return mArr.end(); }

    // serialisation (compact; object keys sorted)
    std::string toString() const;
    static bool parse(const char* begin, const char* end, Value& out, std::string& err);

private:
    void write(std::string& out) const;
    ValueType mType = nullValue;
    int64_t mInt = 0;
    double mReal = 0;
    std::string mStr;
    std::vector<Value> mArr;
    std::map<std::string, Value> mObj;
};

} // namespace Json
