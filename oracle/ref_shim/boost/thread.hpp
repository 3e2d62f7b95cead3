// boost::thread_group / bind / ref stand-in over std::thread (node.cpp:207-217). TEST INFRASTRUCTURE.
#ifndef PAGAN2_B200_SHIM_THREAD_HPP
#define PAGAN2_B200_SHIM_THREAD_HPP
#include <thread>
#include <vector>
#include <functional>
#include "boost/thread/mutex.hpp"
namespace boost {
using std::bind;
using std::ref;
class thread {
public:
    static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
};
class thread_group {
    std::vector<std::thread> members_;
public:
    template <class F> void create_thread(F f) { members_.emplace_back(f); }
    void join_all() { for (auto &t : members_) if (t.joinable()) t.join(); }
    ~thread_group() { join_all(); }
};
}
#endif
