// boost::mutex stand-in over std::mutex (node.h:76-78). TEST INFRASTRUCTURE.
#ifndef PAGAN2_B200_SHIM_THREAD_MUTEX_HPP
#define PAGAN2_B200_SHIM_THREAD_MUTEX_HPP
#include <mutex>
#include <stdexcept>
namespace boost {
class mutex : public std::mutex {
public:
    typedef std::unique_lock<std::mutex> scoped_lock;
};
class lock_error : public std::runtime_error {
public:
    lock_error() : std::runtime_error("lock_error") {}
};
}
#endif
