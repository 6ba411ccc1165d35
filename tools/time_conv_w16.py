import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'time_conv.py')).read().split("B = int(sys.argv[1])")[0])
for S_, ci, co in [(16, 128, 128), (16, 512, 128), (16, 256, 256)]:
    run(64, 32, S_, ci, co)
