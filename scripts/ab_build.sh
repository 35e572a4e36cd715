#!/bin/bash
# build libocean_b200.so variants for on-GPU A/B runs: scripts/ab_build.sh name -DFLAG=.. -DFLAG2=..
# -> gfx_ocean_b200/variants/libocean_b200.<name>.so (own object directory; the default library is not touched);
# select at run time with OCEAN_B200_LIB=<path>
set -e
cd "$(dirname "$0")/.."
name=$1; shift
python -m gfx_ocean_b200.build --variant "$name" "$@" > /dev/null
echo "built variant $name ($*)"
