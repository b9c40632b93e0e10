// TEST INFRASTRUCTURE ONLY (oracle build shim, never linked into the product).
//
// The reference ships src/easyloggingpp/easylogging++.h as a 0-byte placeholder
// that its CMake fetches from GitHub at build time (src/easyloggingpp/CMakeLists.txt:14-31).
// There is no network here, so the oracle build puts this directory first on the
// include path.  Only the three macros the CPU colourer's translation units use
// are provided; every LOG(x) << ... statement compiles to a discarded stream.
#pragma once
#include <ostream>

namespace oracle_shim {
struct NullStream {
	template <typename T> NullStream & operator<<(const T &) { return *this; }
	NullStream & operator<<(std::ostream & (*)(std::ostream &)) { return *this; }
};
}

#define LOG(level) ::oracle_shim::NullStream()
#define INITIALIZE_EASYLOGGINGPP
#define START_EASYLOGGINGPP(argc, argv)
