import os,subprocess,sys
for f in (150,200,250,300,350,400):
    env=dict(os.environ, TKM_MSM_HOST_FIRST=str(f))
    out=subprocess.run([sys.executable,"scripts/e2e_pieces_sweep.py"],env=env,capture_output=True,text=True).stdout.strip().splitlines()
    print(f, [l for l in out if l.startswith(("None","2 ","3 "))])
