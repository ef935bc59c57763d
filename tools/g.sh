#!/bin/bash
# Run a command on the GPU box with gpurun_out/<tag> created first: tools/g.sh <tag> <timeout> '<cmd using $OUT>'
TAG=$1; TMO=$2; shift 2
/usr/local/graft/bin/gpurun --timeout $TMO -- "export OUT=gpurun_out/$TAG; mkdir -p \$OUT; $*"
