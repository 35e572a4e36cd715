mkdir -p gpurun_out
B="python bench.py --steps 1000 --no-cpu-baseline --no-extras"
V=$PWD/gfx_ocean_b200/variants
run() { name=$1; shift; env "$@" timeout 300 $B $EXTRA > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; }
EXTRA=""
run tma OCEAN_B200_ROWS=tma
for v in pf370 pf740 pf1480 m6; do run tma_$v OCEAN_B200_ROWS=tma OCEAN_B200_LIB=$V/libocean_b200.$v.so; done
EXTRA="--tiles 16"
run t16_tma OCEAN_B200_ROWS=tma
run t16_tma_pf740 OCEAN_B200_ROWS=tma OCEAN_B200_LIB=$V/libocean_b200.pf740.so
