// minilua command-line runner: `minilua script.lua` or `minilua -e "code"`.
// Used by the CPU test-suite to exercise the evaluator like a `lua` binary.
#include <cstdio>
#include <cstring>
#include <string>

#include "minilua.h"

int main(int argc, char **argv) {
    minilua::State L;
    // the aliases the fisheye layer injects (engine/NQ/fisheye.c:1230-1248) are
    // NOT added here: this is a plain interpreter.
    try {
        for (int i = 1; i < argc; ++i) {
            if (!strcmp(argv[i], "-e") && i + 1 < argc) {
                L.run(argv[++i], "=(command line)");
            } else {
                minilua::Value fn = L.load_file(argv[i]);
                minilua::ValueList out;
                L.call(fn, nullptr, 0, out);
            }
        }
    } catch (minilua::LuaError &e) {
        fprintf(stderr, "minilua: %s\n", e.what());
        return 1;
    }
    return 0;
}
