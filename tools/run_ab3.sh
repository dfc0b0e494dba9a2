set -x
A=$PWD/build/libA.so; B=$PWD/build/libB.so
for cfg in "8 4" "8 0"; do
  CVXPNPL_B200_LIB=$A python tools/ab_step.py $cfg
  CVXPNPL_B200_LIB=$B python tools/ab_step.py $cfg
  CVXPNPL_B200_LIB=$B CVXPNPL_B200_SVC_ITERS=100000 python tools/ab_step.py $cfg
done
CVXPNPL_B200_LIB=$A ADMM=f32 python tools/ab_step.py 0 6
CVXPNPL_B200_LIB=$B ADMM=f32 python tools/ab_step.py 0 6
CVXPNPL_B200_LIB=$B ADMM=f32 CVXPNPL_B200_SVC_ITERS=100000 python tools/ab_step.py 0 6
for cfg in "3 0 1667" "0 3 1667" "4 0 1667"; do
  CVXPNPL_B200_LIB=$A NOISE=0 python tools/ab_step.py $cfg
  CVXPNPL_B200_LIB=$B NOISE=0 python tools/ab_step.py $cfg
done
CVXPNPL_B200_LIB=$A python tools/ab_step.py 8 4 12500
CVXPNPL_B200_LIB=$B python tools/ab_step.py 8 4 12500
