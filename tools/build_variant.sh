#!/bin/bash
# build_variant.sh NAME [nvcc -D flags...] : builds lua-multigrid-poisson_b200/libmgpoisson_NAME.so (experiments; the
# default library is built by `make` in csrc/ or __graft_entry__.build()).
name=$1; shift
cd "$(dirname "$0")/../lua-multigrid-poisson_b200/csrc" && make EXTRA="$*" OUT=../libmgpoisson_$name.so LOG=/tmp/build_$name.log > /tmp/make_$name.log 2>&1
echo "$name rc=$?"
