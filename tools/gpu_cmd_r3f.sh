for tr in 2 30; do echo "== TRACE=$tr"; B200LDLT_TRACE=$tr timeout 300 python tools/profile_run.py 64 2 2>&1 | grep "k_front" | grep "m=291" | tail -1; done
