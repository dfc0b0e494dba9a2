C=$PWD/build/libC.so
export CVXPNPL_B200_LIB=$C
for cfg in "6 0" "10 0" "12 0" "0 8" "5 3" "6 6" "4 4"; do for n in 2 4 6; do
  echo "== cfg $cfg SVC_CTAS=$n"
  CVXPNPL_B200_SVC_CTAS=$n python tools/ab_step.py $cfg
done; done
for n in 3 6 8; do
  echo "== cfg 8 0 SVC_CTAS=$n"
  CVXPNPL_B200_SVC_CTAS=$n python tools/ab_step.py 8 0
done
for n in 3; do
  echo "== cfg 8 4 SVC_CTAS=$n"
  CVXPNPL_B200_SVC_CTAS=$n python tools/ab_step.py 8 4
done
