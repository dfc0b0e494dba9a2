C=$PWD/build/libC.so
export CVXPNPL_B200_LIB=$C
for n in 2 4 8 16; do for it in 400 100000; do
  echo "== SVC_CTAS=$n SVC_ITERS=$it"
  CVXPNPL_B200_SVC_CTAS=$n CVXPNPL_B200_SVC_ITERS=$it ADMM=f32 python tools/ab_step.py 0 6
done; done
for n in 2 8; do for it in 400 100000; do
  echo "== SVC_CTAS=$n SVC_ITERS=$it"
  CVXPNPL_B200_SVC_CTAS=$n CVXPNPL_B200_SVC_ITERS=$it python tools/ab_step.py 4 0
  CVXPNPL_B200_SVC_CTAS=$n CVXPNPL_B200_SVC_ITERS=$it python tools/ab_step.py 0 6
done; done
for n in 2 4; do for it in 400 100000; do
  echo "== SVC_CTAS=$n SVC_ITERS=$it"
  CVXPNPL_B200_SVC_CTAS=$n CVXPNPL_B200_SVC_ITERS=$it python tools/ab_step.py 8 0
  CVXPNPL_B200_SVC_CTAS=$n CVXPNPL_B200_SVC_ITERS=$it python tools/ab_step.py 8 4
done; done
