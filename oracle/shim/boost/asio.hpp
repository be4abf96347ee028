// Shim: just enough of boost::asio / boost::thread / boost::bind for scoring_function/score_calculator.cpp and the drivers'
// main() functions.  The per-variable time limit (-r) is not exercised through oracle/_ref: async_wait never fires.
#pragma once
#include <functional>
#include <thread>
namespace boost {
namespace system { class error_code {}; }
namespace posix_time { struct seconds { explicit seconds(long) {} }; }
namespace asio {
class io_service { public: void run() {} void stop() {} void reset() {} };
namespace placeholders { static const int error = 0; }
class deadline_timer {
public:
    explicit deadline_timer(io_service &) {}
    template <class T> void expires_from_now(const T &) {}
    template <class F> void async_wait(F) {}
    void cancel() {}
};
} // namespace asio
template <class T> inline std::reference_wrapper<T> ref(T &t) { return std::ref(t); }
struct bound_nothing { void operator()() const {} };
template <class... A> inline bound_nothing bind(A &&...) { return bound_nothing(); }
class thread { // a real thread (score_main.cpp's workers must run); detached on destruction like boost::thread
public:
    template <class F, class... A> explicit thread(F f, A... a) : t(f, a...) {}
    ~thread() { if (t.joinable()) t.detach(); }
    void join() { if (t.joinable()) t.join(); }
private:
    std::thread t;
};
} // namespace boost
