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
  auto named = coregex::Regex::Compile("(?P<year>\\d+)-(?P<month>\\d+)");
  if (named.SubexpNames() != std::vector<std::string>{"", "year", "month"} || named.SubexpIndex("month") != 2 ||
      named.SubexpIndex("day") != -1) return 7;
  if (coregex::QuoteMeta("hello.world") != "hello\\.world") return 8;
  std::string ex;
  const int64_t m123[4] = {5, 8, 5, 8};
  coregex::Regex::Expand(ex, "[$0|$1|$$|${1}|$9|$", "test 123 end", m123, 4);   // replace_test.go:197-226
  if (ex != "[123|123|$|${1}||$") return 9;
  cgx_config cfg;
  cgx_default_config(&cfg);
  auto rc = coregex::Regex::CompileWithConfig("a+", cfg);
  rc.Longest();
  const char* h = "1.2 3.4";
  try {
    auto m = re.FindAllIndex((const uint8_t*)h, 7);
    std::printf("matches %zu\n", m.size());      // GPU present
    if (m.size() != 2 || m[0].first != 0 || m[0].second != 3) return 6;
    // replace_test.go:75-104, :159-195
    auto email = coregex::Regex::Compile("(\\w+)@(\\w+)\\.(\\w+)");
    if (email.ReplaceAllString("user@example.com", "$1 at $2 dot $3") != "user at example dot com") return 10;
    auto num = coregex::Regex::Compile("\\d+");
    if (num.ReplaceAllString("age: 42", "[$0]") != "age: [42]" || num.ReplaceAllLiteralString("1 2 3", "X") != "X X X") return 11;
    if (num.ReplaceAllStringFunc("1 2 3", [](const std::string& s) { return std::to_string(2 * std::stoi(s)); }) != "2 4 6") return 12;
    auto comma = coregex::Regex::Compile(",");
    if (comma.Split("a,b,c,d,e", 3) != std::vector<std::string>{"a", "b", "c,d,e"} || !comma.Split("a,b", 0).empty()) return 13;
    if (coregex::Regex::Compile("a").Split("aaa") != std::vector<std::string>{"", "", "", ""}) return 14;
    if (num.FindString("age: 42") != "42" || num.FindStringIndex("none").first != -1 ||
        num.FindAllString("1 22 333") != std::vector<std::string>{"1", "22", "333"}) return 15;
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


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cpp_header_wrappers_on_device():
    """The same program on a GPU box: the search branch runs (FindAllIndex, ReplaceAll*, Split, Find*
    through the C++ mirror) and prints the match count instead of the no-device message."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(d, "t")
        lib = os.path.join(ROOT, "coregex_b200", "lib")
        subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-L", lib,
                               "-lcoregex_b200", "-Wl,-rpath," + lib, "-o", exe])
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "matches 2" in r.stdout, (r.returncode, r.stdout, r.stderr)
