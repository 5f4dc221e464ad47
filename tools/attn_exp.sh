for d in 0 1 2 4 6 7; do echo "DBG=$d"; SPE_ATTN_DBG=$d python tools/attn_micro.py 2>&1 | sed -n 2,2p; done
