for B in 1048576 131072 32768; do for t in 28 36 44; do echo "B=$B tail=$t: $(B200MPC_HAND_ITER_TAIL=$t VAR=B KIND=lane B=$B NREP=3 python tools/prof_solve.py 2>&1 | sed -n 3p)"; done; done
for t in 0 48 64 96; do echo "A 262144 tail=$t: $(B200MPC_HAND_ITER_TAIL=$t VAR=A KIND=lane B=262144 NREP=2 python tools/prof_solve.py 2>&1 | sed -n 2p)"; done
