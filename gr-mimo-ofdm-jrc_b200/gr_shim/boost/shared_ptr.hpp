// boost/shared_ptr.hpp stand-in: GNU Radio 3.8 block sptrs are boost::shared_ptr; without Boost the
// shim maps them onto the standard library (same semantics for everything the blocks use).
#ifndef JRC_SHIM_BOOST_SHARED_PTR_HPP
#define JRC_SHIM_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::enable_shared_from_this;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::make_shared;
}  // namespace boost
#endif
