#!/bin/bash
# build in-tree, then run a command on the GPU box:  tools/g.sh [gpurun flags --] 'command'
set -e
make -C /root/repo/mir_prefer_b200/csrc -j6 2>&1 | grep -E "error|Error" && exit 1
make -C /root/repo/oracle all >/dev/null
exec /usr/local/graft/bin/gpurun --timeout ${GTIMEOUT:-1500} -- "$@"
