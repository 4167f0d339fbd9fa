// roundtrip_main.cpp -- the shape of zpaq_test.v:364-384 and cmd/main.v:288-401 over the C++ mirror.
// Built by tests/test_gpu_cpp_mirror.py on the GPU box:  g++ -Iinclude -Izpaq-v_b200/host ... -lzpaqgpu
#include <cstdio>
#include <cstring>

#include "zpaq_gpu.hpp"

int main(int argc, char **argv) {
    const int level = argc > 1 ? std::atoi(argv[1]) : 1;
    std::vector<uint8_t> input;
    for (int c; (c = std::getchar()) != EOF;) input.push_back(uint8_t(c));
    zpaq::FileReader in(input);
    zpaq::FileWriter out;
    zpaq::Compressor comp;
    comp.set_input(&in);
    comp.set_output(&out);
    comp.start_block(level);
    comp.start_segment("test", "");
    while (comp.compress(65536)) {}
    comp.end_segment();
    comp.end_block();
    zpaq::FileReader arc(out.bytes());
    zpaq::FileWriter back;
    zpaq::Decompresser d;
    d.set_input(&arc);
    d.set_output(&back);
    int segs = 0, ok = 0;
    while (d.find_block())
        while (d.find_filename()) {
            while (d.decompress(65536)) {}
            d.read_segment_end();
            ++segs, ok += d.last_sha1_ok() == 1;
        }
    const bool same = back.bytes() == input;
    for (uint8_t b : out.bytes()) std::printf("%02x", b);
    std::printf("\n%d %d %d\n", segs, ok, same ? 1 : 0);
    // JidacArchive.create_archive (jidac.v:181) of two files made of the input: printed as a third line
    zpaq::FileWriter jw;
    zpaq::JidacArchive ja(20260101120000LL);
    ja.set_output(&jw);
    ja.create_archive({{"a", input}, {"b", std::vector<uint8_t>(input.rbegin(), input.rend())}}, 0);
    for (uint8_t b : jw.bytes()) std::printf("%02x", b);
    std::printf("\n");
    // batch mode: two blocks queued, one launch at flush(); the Writer must hold the single block twice
    zpaq::FileWriter bw;
    zpaq::Compressor batch(true);
    batch.set_output(&bw);
    for (int k = 0; k < 2; ++k) {
        zpaq::FileReader again(input);
        batch.set_input(&again);
        batch.start_block(level);
        batch.start_segment("test", "");
        while (batch.compress(65536)) {}
        batch.end_segment();
        batch.end_block();
    }
    const bool nothing_yet = bw.bytes().empty();
    batch.flush();
    std::vector<uint8_t> twice = out.bytes();
    twice.insert(twice.end(), out.bytes().begin(), out.bytes().end());
    const bool batch_ok = nothing_yet && bw.bytes() == twice;
    std::printf("batch %d\n", batch_ok ? 1 : 0);
    return same && segs == 1 && ok == 1 && batch_ok ? 0 : 1;
}
