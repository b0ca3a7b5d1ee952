for q in 0 1; do echo "queue=$q"; VEL_SYRK_QUEUE=$q timeout 300 python tools/syrk_sweep.py 2 4 5 6 8 10 12 16 2>&1 | tail -8; done
