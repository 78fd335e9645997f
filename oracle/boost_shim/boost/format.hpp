// Minimal stand-in for <boost/format.hpp>: positional "%N%" substitution only, which is all the
// reference uses (src/external_sort.hpp:107, src/file_utils.cpp:100,122). Oracle build only.
#pragma once
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
public:
    explicit format(const char* f) : m_fmt(f) {}
    explicit format(const std::string& f) : m_fmt(f) {}
    template <class T>
    format& operator%(const T& v) {
        std::ostringstream os;
        os << v;
        m_args.push_back(os.str());
        return *this;
    }
    std::string str() const {
        std::string out;
        for (size_t i = 0; i < m_fmt.size(); ++i) {
            if (m_fmt[i] == '%') {
                size_t j = i + 1, n = 0;
                while (j < m_fmt.size() && m_fmt[j] >= '0' && m_fmt[j] <= '9') { n = n * 10 + (m_fmt[j] - '0'); ++j; }
                if (j < m_fmt.size() && m_fmt[j] == '%' && j > i + 1 && n >= 1 && n <= m_args.size()) {
                    out += m_args[n - 1];
                    i = j;
                    continue;
                }
            }
            out += m_fmt[i];
        }
        return out;
    }
private:
    std::string m_fmt;
    std::vector<std::string> m_args;
};
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}  // namespace boost
