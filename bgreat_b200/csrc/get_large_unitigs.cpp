// getLargeUnitigs -- drop-in for the reference's helper of the same name (reference getLargeUnitigs.cpp:43-74,
// the hand-off of the graph to a read aligner for the reads BGREAT leaves unmapped):
//   getLargeUnitigs IN OUT LENGTH
// writes ONE FASTA record ">N" whose sequence is the concatenation of every line of IN that is longer than
// LENGTH, each without its last character (the ';' of the unitig format the tool was written for) and in
// upper case. Kept exactly as the reference behaves, which includes: header lines count as lines; LENGTH is
// compared as an unsigned number (a negative LENGTH selects nothing); nothing is printed for a wrong argument
// count except the usage line when there are fewer than three arguments. One difference: an input that cannot
// be opened gives the empty record (the reference loops forever on `while(!in.eof())` of a failed stream).
// Host only; the input is streamed through a large buffer, lines may be arbitrarily long.
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

static void get_big_unitigs(const char* input, const char* output, int length) {
    FILE* in = fopen(input, "rb");
    FILE* out = fopen(output, "wb");  // truncated even when the input is missing, like the reference's ofstream
    if (!out) {
        if (in) fclose(in);
        return;
    }
    fputs(">N\n", out);
    const size_t min_size = (size_t)length;  // `line.size() > length`: the int is converted to size_t
    std::string line;
    std::vector<char> buf(1 << 20);
    auto flush_line = [&]() {
        if (line.size() > min_size) {
            line.pop_back();
            for (char& c : line) c = (char)toupper((unsigned char)c);
            fwrite(line.data(), 1, line.size(), out);
        }
        line.clear();
    };
    if (in) {
        size_t got;
        while ((got = fread(buf.data(), 1, buf.size(), in)) > 0) {
            size_t start = 0;
            for (size_t i = 0; i < got; ++i)
                if (buf[i] == '\n') {
                    line.append(buf.data() + start, i - start);
                    flush_line();
                    start = i + 1;
                }
            line.append(buf.data() + start, got - start);
        }
        flush_line();  // last line without a newline (an empty one selects nothing: size 0 is never > LENGTH)
        fclose(in);
    }
    fputc('\n', out);
    fclose(out);
}

int main(int argc, char** argv) {
    if (argc < 4) printf("Need in, out and length...\n");
    if (argc == 4) get_big_unitigs(argv[1], argv[2], atoi(argv[3]));
    return 0;
}
