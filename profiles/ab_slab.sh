L=fly-by-cnn_b200/flyby_b200/lib/libflyby_b200.so
for px in 67108864 12582912 8388608 6291456 4194304 2097152; do echo "== FLYBY_SLAB_PIXELS=$px"; FLYBY_SLAB_PIXELS=$px python profiles/ab_render.py --one $L 15; done
