C=$PWD/build/libC.so; D=$PWD/build/libD.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2be_pytest.txt; tail -3 gpurun_out/r2be_pytest.txt
for cfg in "3 0 1667" "0 3 1667" "4 0 1667"; do
  CVXPNPL_B200_LIB=$C NOISE=0 python tools/ab_step.py $cfg
  CVXPNPL_B200_LIB=$D NOISE=0 python tools/ab_step.py $cfg
done
CVXPNPL_B200_LIB=$C ADMM=f32 python tools/ab_step.py 0 6
CVXPNPL_B200_LIB=$D ADMM=f32 python tools/ab_step.py 0 6
for cfg in "8 0" "8 4" "0 8" "5 3" "8 4 12500"; do
  CVXPNPL_B200_LIB=$D python tools/ab_step.py $cfg
done
