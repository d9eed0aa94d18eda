for caps in 0 "64,160" "48,128" "96" "64,128,224"; do echo A caps=$caps; B200MPC_WARP_CAPS=$caps VAR=A KIND=warp B=4096 NREP=3 python tools/prof_solve.py 2>&1 | sed -n 2,3p; done
for caps in 0 32 24 "24,48"; do echo B caps=$caps; B200MPC_WARP_CAPS=$caps VAR=B KIND=warp B=4096 NREP=3 python tools/prof_solve.py 2>&1 | sed -n 2,3p; done
for caps in 0 32; do echo B16k caps=$caps; B200MPC_WARP_CAPS=$caps VAR=B KIND=warp B=16384 NREP=3 python tools/prof_solve.py 2>&1 | sed -n 3p; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
