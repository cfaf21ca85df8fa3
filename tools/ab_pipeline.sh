for pl in 0 1; do for mb in 8 7; do echo "PIPELINE=$pl MINBLOCKS=$mb"; MDPP_JIT_PIPELINE=$pl MDPP_JIT_MINBLOCKS=$mb python tools/time_headline.py 2>&1 | grep "C2 fast\|C1"; done; done
