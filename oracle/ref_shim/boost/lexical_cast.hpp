// TEST INFRASTRUCTURE ONLY.  Stand-in for <boost/lexical_cast.hpp> (render helpers only, not on the path).
#ifndef VSO_REF_SHIM_BOOST_LEXICAL_CAST_HPP_
#define VSO_REF_SHIM_BOOST_LEXICAL_CAST_HPP_
#include <sstream>
#include <string>
namespace boost {
template <class To, class From> To lexical_cast(const From& f) {
  std::stringstream s;
  s << f;
  To t;
  s >> t;
  return t;
}
}  // namespace boost
#endif
