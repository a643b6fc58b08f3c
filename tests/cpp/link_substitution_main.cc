// main() of test_link_substitution: runs the reference's own unit tests (registered through the gtest stand-in) against
// liblzgpu.so.  Usage: test_link_substitution [substring filter]   ("-Benchmark" as the first argument skips the benchmarks).
// Prints one line per test; exit code = number of failed expectations (0 = all passed).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <gtest/gtest.h>

int main(int argc, char **argv) {
	const char *only = nullptr, *skip = nullptr;
	for (int i = 1; i < argc; ++i) {
		if (argv[i][0] == '-') skip = argv[i] + 1;
		else only = argv[i];
	}
	int ran = 0;
	for (auto &t : ::testing::registry()) {
		const std::string full = t.suite + "." + t.name;
		if (only && full.find(only) == std::string::npos) continue;
		if (skip && full.find(skip) != std::string::npos) continue;
		const int before = ::testing::failures();
		const auto t0 = std::chrono::steady_clock::now();
		t.body();
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		std::printf("[%s] %s (%.1f ms)\n", ::testing::failures() == before ? "  OK  " : "FAILED", full.c_str(), ms);
		std::fflush(stdout);
		++ran;
	}
	std::printf("%d tests, %d failed expectations\n", ran, ::testing::failures());
	return ::testing::failures() > 255 ? 255 : ::testing::failures();
}
