#!/bin/bash
# build libocean_b200.so variants for on-GPU A/B runs: scripts/ab_build.sh name "-DFLAG=.. -DFLAG2=.."
# -> gfx_ocean_b200/variants/libocean_b200.<name>.so ; select at run time with OCEAN_B200_LIB=<path>
set -e
cd "$(dirname "$0")/.."
mkdir -p gfx_ocean_b200/variants
OCEAN_NVCC_EXTRA="$2" python -m gfx_ocean_b200.build --force > /dev/null
cp gfx_ocean_b200/libocean_b200.so gfx_ocean_b200/variants/libocean_b200.$1.so
echo "built variant $1 ($2)"
