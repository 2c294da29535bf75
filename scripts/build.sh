#!/bin/bash
# build the CUDA library; non-zero exit (and the compiler's messages) on failure
make -C "$(dirname "$0")/../panda_lang_manip_b200/csrc" -j8 > /tmp/pg_build.log 2>&1 || { grep -E "error|Error" /tmp/pg_build.log | head -20; exit 1; }
echo "build ok"
