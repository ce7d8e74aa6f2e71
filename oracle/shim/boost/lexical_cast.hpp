// oracle/shim/boost/lexical_cast.hpp -- stand-in for the one use in the reference
// (scr/subset_to_test_and_training.cpp: std::transform(..., boost::lexical_cast<int, std::string>)).
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : public std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <class Target, class Source>
inline Target lexical_cast(const Source& s) {
    std::stringstream ss;
    ss << s;
    Target t;
    ss >> t;
    if (ss.fail() || !(ss >> std::ws).eof()) throw bad_lexical_cast();
    return t;
}
}  // namespace boost
