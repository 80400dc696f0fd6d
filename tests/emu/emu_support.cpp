// Emulator support entry points (test infrastructure).
#include <condition_variable>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

// 1 when this process may run `n` more OS threads at the same time (thread-block cluster emulation needs 8 x 512 of them): the threads are
// created, parked on a condition variable and released again; a std::system_error from std::thread (pid / rlimit exhaustion) -> 0.
extern "C" int emu_can_spawn(int n) {
  std::mutex mu;
  std::condition_variable cv;
  bool go = false;
  std::vector<std::thread> ths;
  bool ok = true;
  try {
    ths.reserve(n);
    for (int i = 0; i < n; ++i)
      ths.emplace_back([&] {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return go; });
      });
  } catch (const std::system_error&) {
    ok = false;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    go = true;
  }
  cv.notify_all();
  for (auto& t : ths) t.join();
  return ok ? 1 : 0;
}
