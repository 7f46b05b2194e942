// TEST INFRASTRUCTURE: stand-in for boost::get_system_time.
#pragma once
#include <boost/date_time/posix_time/posix_time.hpp>
namespace boost { inline posix_time::ptime get_system_time() { return posix_time::microsec_clock::local_time(); } }
