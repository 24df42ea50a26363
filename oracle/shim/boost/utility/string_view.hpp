// Minimal stand-in for <boost/utility/string_view.hpp> (Boost is not installable in this image): just enough of
// boost::string_view for the reference translation units compiled IN PLACE by oracle/build_ref.sh.
// Written for this repo; not Boost code.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <ostream>
#include <string>
#include <string_view>

namespace boost {
class string_view : public std::string_view {
public:
    using std::string_view::string_view;
    constexpr string_view() noexcept = default;
    constexpr string_view(const std::string_view& v) noexcept : std::string_view(v) {}
    string_view(const std::string& s) noexcept : std::string_view(s) {}
    std::string to_string() const { return std::string(data(), size()); }
    string_view substr(size_t pos = 0, size_t n = npos) const { return string_view(std::string_view::substr(pos, n)); }
    bool starts_with(string_view x) const { return size() >= x.size() && compare(0, x.size(), x) == 0; }
    bool ends_with(string_view x) const { return size() >= x.size() && compare(size() - x.size(), x.size(), x) == 0; }
};
} // namespace boost

namespace std {
template <>
struct hash<boost::string_view> {
    size_t operator()(const boost::string_view& v) const noexcept { return hash<std::string_view>()(v); }
};
} // namespace std
