"""The C++ mirror (include/coregex.hpp) compiles against the C ABI and behaves like the Go API on
the CPU tier: Compile succeeds, syntax errors carry the stdlib text, searches without a device
raise (no fallback)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cstdio>
#include <cstring>
#include "coregex.hpp"
int main() {
  auto re = coregex::Regex::Compile("(\\d+)\\.(\\d+)");
  if (re.NumSubexp() != 2 || re.String() != "(\\d+)\\.(\\d+)") return 2;
  if (re.Strategy().empty()) return 3;
  try { coregex::Regex::MustCompile("a**"); return 4; }
  catch (const coregex::Error& e) {
    if (std::strcmp(e.what(), "regexp: Compile(`a**`): error parsing regexp: invalid nested repetition operator: `**`")) return 5;
  }
  const char* h = "1.2 3.4";
  try {
    auto m = re.FindAllIndex((const uint8_t*)h, 7);
    std::printf("matches %zu\n", m.size());      // GPU present
    if (m.size() != 2 || m[0].first != 0 || m[0].second != 3) return 6;
  } catch (const coregex::Error& e) {
    std::printf("no device: %s\n", e.what());    // CPU box: must fail loudly, not fall back
  }
  return 0;
}
'''


def test_cpp_header_compiles_and_runs():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(d, "t")
        lib = os.path.join(ROOT, "coregex_b200", "lib")
        subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-L", lib,
                               "-lcoregex_b200", "-Wl,-rpath," + lib, "-o", exe])
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
