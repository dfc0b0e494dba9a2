E=$PWD/build/libE.so
for cfg in "3 0 1667" "4 0 1667"; do
  CVXPNPL_B200_LIB=$E NOISE=0 python tools/ab_step.py $cfg
  CVXPNPL_B200_LIB=$E NOISE=0 CVXPNPL_B200_QUAD_BUDGET=0 python tools/ab_step.py $cfg
done
CVXPNPL_B200_LIB=$E ADMM=f32 python tools/ab_step.py 0 6
CVXPNPL_B200_LIB=$E ADMM=f32 CVXPNPL_B200_QUAD_BUDGET=0 python tools/ab_step.py 0 6
CVXPNPL_B200_LIB=$E python tools/ab_step.py 8 4
CVXPNPL_B200_LIB=$E python tools/ab_step.py 8 0
