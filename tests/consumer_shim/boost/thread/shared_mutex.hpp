// stand-in for <boost/thread/shared_mutex.hpp> + locks: boost::shared_mutex / shared_lock / unique_lock on the C++14 primitives
#pragma once
#include <mutex>
#include <shared_mutex>
namespace boost {
typedef std::shared_timed_mutex shared_mutex;
template <typename M> using shared_lock = std::shared_lock<M>;
template <typename M> using unique_lock = std::unique_lock<M>;
}  // namespace boost
