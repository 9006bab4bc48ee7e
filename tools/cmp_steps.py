import subprocess, sys
a=subprocess.run(["python","tools/steps.py",sys.argv[1]],capture_output=True,text=True).stdout.splitlines()
b=subprocess.run(["python","tools/steps.py",sys.argv[2]],capture_output=True,text=True).stdout.splitlines()
for x,y in zip(a,b):
    print(x[:44], "|", y[24:44])
