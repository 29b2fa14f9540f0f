// Stand-in for MSVC's <ppl.h>: the reference fans tiles (Renderer.hpp:75-433) and framebuffer tiles (:438-476) out with
// concurrency::parallel_for. Tiles are independent (each writes its own accumulator tile / framebuffer block), so any schedule computes
// the same values; this one hands indices to REF_THREADS worker threads (default: all hardware threads) from an atomic counter.
// TEST INFRASTRUCTURE ONLY (oracle/ref_renderer_build.sh).
#pragma once
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>
namespace concurrency {
struct auto_partitioner {};
inline unsigned ref_thread_count() {
	if (const char* e = std::getenv("REF_THREADS")) { const int n = std::atoi(e); if (n > 0) return static_cast<unsigned>(n); }
	const unsigned n = std::thread::hardware_concurrency();
	return n ? n : 1u;
}
template <class Index, class F> void parallel_for(Index first, Index last, F&& f) {
	const unsigned threads = ref_thread_count();
	if (threads <= 1 || last - first < 2) { for (Index i = first; i < last; ++i) f(i); return; }
	std::atomic<Index> next{first};
	auto work = [&] { for (;;) { const Index i = next.fetch_add(1); if (i >= last) break; f(i); } };
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < threads; t++) pool.emplace_back(work);
	work();
	for (auto& t : pool) t.join();
}
template <class Index, class F> void parallel_for(Index first, Index last, F&& f, const auto_partitioner&) { parallel_for(first, last, static_cast<F&&>(f)); }
}  // namespace concurrency
