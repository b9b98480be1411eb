// dictionary-prep: the reference's dictionary preprocessing tool (src/runner/dictionary-prep.cpp), same command line:
//   dictionary-prep -e dictionary input output     encode
//   dictionary-prep -d dictionary input output     decode
// Host-only (gmix_b200/host/dictionary.h); byte-compatible with the reference's tool in both directions.
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <fstream>
#include <iterator>
#include <vector>

#include "dictionary.h"

static bool ReadAll(const char* path, std::vector<uint8_t>* v) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  v->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return true;
}

static int Help() {
  printf("This tool runs dictionary preprocessing.\n");
  printf("Encode: dictionary-prep -e [dictionary] [input] [output]\n");
  printf("Decode: dictionary-prep -d [dictionary] [input] [output]\n");
  return -1;
}

int main(int argc, char* argv[]) {
  if (argc != 5 || strlen(argv[1]) != 2 || argv[1][0] != '-' || (argv[1][1] != 'e' && argv[1][1] != 'd')) return Help();
  const clock_t start = clock();
  std::vector<uint8_t> dict, in;
  if (!ReadAll(argv[2], &dict) || !ReadAll(argv[3], &in)) return Help();
  const gmixb::WordTransform t(dict);
  const std::vector<uint8_t> out = argv[1][1] == 'e' ? t.Encode(in) : t.Decode(in);
  std::ofstream f(argv[4], std::ios::binary);
  if (!f.is_open()) return Help();
  f.write((const char*)out.data(), (std::streamsize)out.size());
  printf("\r%zu bytes -> %zu bytes in %1.2f s.\n", in.size(), out.size(), ((double)clock() - start) / CLOCKS_PER_SEC);
  return f.good() ? 0 : -1;
}
