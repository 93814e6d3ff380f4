// boost::split / boost::is_any_of, as simulator_util.cpp uses them.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of { std::string set; explicit is_any_of(const std::string& s) : set(s) {} };
template <typename C> void split(C& out, const std::string& in, const is_any_of& pred) {
    out.clear();
    std::string cur;
    for (char ch : in) {
        if (pred.set.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur.push_back(ch);
    }
    out.push_back(cur);
}
}  // namespace boost
