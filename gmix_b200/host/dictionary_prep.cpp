// dictionary-prep: command line of the reference's dictionary preprocessing tool (src/runner/dictionary-prep.cpp) over
// gmixb::WordTransform (dictionary.h). Host-only; files are interchangeable with the reference tool's in both directions.
//   dictionary-prep -e <dictionary> <input> <output>     text -> word codes
//   dictionary-prep -d <dictionary> <input> <output>     word codes -> text
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "dictionary.h"

namespace {

using Bytes = std::vector<uint8_t>;

bool Slurp(const std::string& path, Bytes* bytes) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  bytes->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return true;
}

int Usage() {
  std::fputs("This tool runs dictionary preprocessing.\n"
             "Encode: dictionary-prep -e [dictionary] [input] [output]\n"
             "Decode: dictionary-prep -d [dictionary] [input] [output]\n", stdout);
  return -1;
}

}  // namespace

int main(int argc, char* argv[]) {
  const std::string mode = argc == 5 ? argv[1] : "";
  if (mode != "-e" && mode != "-d") return Usage();
  const auto t0 = std::chrono::steady_clock::now();
  Bytes dictionary, input;
  if (!Slurp(argv[2], &dictionary) || !Slurp(argv[3], &input)) return Usage();
  const gmixb::WordTransform transform(dictionary);
  const Bytes output = mode == "-e" ? transform.Encode(input) : transform.Decode(input);
  std::ofstream sink(argv[4], std::ios::binary);
  if (!sink) return Usage();
  sink.write(reinterpret_cast<const char*>(output.data()), static_cast<std::streamsize>(output.size()));
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("\r%zu bytes -> %zu bytes in %1.2f s.\n", input.size(), output.size(), seconds);
  return sink.good() ? 0 : -1;
}
