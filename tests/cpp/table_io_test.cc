// Host-only driver of the table I/O (kaldi-aslp_b200/host/table.h): copies a float-matrix table from an rspecifier to a wspecifier,
// like copy-feats does, so that tests/test_cpu_table_io.py can push archives through files, pipes, stdin / stdout and compressed
// matrices without a device.
#include <iostream>
#include "table.h"

int main(int argc, char** argv) {
  using namespace kaldi;
  try {
    if (argc != 3) { std::cerr << "usage: table_io_test <rspecifier> <wspecifier>\n"; return 2; }
    SequentialBaseFloatMatrixReader reader(argv[1]);
    BaseFloatMatrixWriter writer(argv[2]);
    int n = 0;
    for (; !reader.Done(); reader.Next(), ++n) writer.Write(reader.Key(), reader.Value());
    std::cerr << "copied " << n << " matrices\n";
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
}
