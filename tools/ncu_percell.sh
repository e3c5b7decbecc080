# ncu capture of the per-cell K_fit on the configs[3] shape (250k cells, 12 coefficients, size factors): one wave of gene tiles
mkdir -p gpurun_out
export DXB200_CFG4_GENES=${DXB200_CFG4_GENES:-9472}
export DXB200_CFG4_CELLS=${DXB200_CFG4_CELLS:-40000}   # the profile replays the kernel ~40 times: a fifth of the cells, same tiles
python tools/scale_check.py cfg4 > gpurun_out/r2_percell_plain.json 2> gpurun_out/r2_percell_plain.err || exit 1
tail -1 gpurun_out/r2_percell_plain.json
ncu --set full --clock-control none --import-source on -k regex:dx_fit_tile --launch-skip 1 -c 1 -o gpurun_out/r2_percell -f python tools/scale_check.py cfg4 > gpurun_out/r2_percell_ncu.log 2>&1
tail -2 gpurun_out/r2_percell_ncu.log
