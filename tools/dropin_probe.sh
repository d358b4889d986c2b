set -e
cd $GRAFT_REPO_ROOT
python - <<'PY'
import os,sys,subprocess
sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import make_golden as mg
R='oracle/_ref'; B='libxaac_b200/dropin/_build/xaacdec_b200'
os.makedirs('/tmp/dp',exist_ok=True)
def run(cmd,env=None):
    r=subprocess.run(cmd,stdout=subprocess.PIPE,stderr=subprocess.PIPE,env=env); return r.stderr.decode(errors='replace')
cases=[('lc',44100,2,12.0,['-aot:2','-adts:1','-br:128000'],False),('lcm',44100,1,12.0,['-aot:2','-adts:1','-br:64000'],False),('usac',32000,2,20.0,['-aot:42','-br:64000','-ccfl_idx:3'],True),('usach',32000,2,20.0,['-aot:42','-br:64000','-ccfl_idx:3','-harmonic_sbr:1'],True),
       ('v1s',48000,2,12.0,['-aot:5','-adts:1','-br:48000'],False),('v2',44100,2,12.0,['-aot:29','-adts:1','-br:32000'],False)]
for name,fs,ch,secs,enc,mp4 in cases:
    wav=f'/tmp/dp/{name}.wav'; mg.write_wav(wav,mg.synth(fs,secs,ch,7),fs)
    bits=f'/tmp/dp/{name}.'+('mp4' if mp4 else 'aac')
    run([R+'/xaacenc',f'-ifile:{wav}',f'-ofile:{bits}']+enc)
    extra=[f'-imeta:/tmp/dp/{name}.txt','-mp4:1'] if mp4 else []
    run([R+'/xaacdec',f'-ifile:{bits}',f'-ofile:/tmp/dp/{name}_ref.wav']+extra)
    log=run([B,f'-ifile:{bits}',f'-ofile:/tmp/dp/{name}_b.wav']+extra,env=dict(os.environ,IXHEAACD_B200_STATS='2'))
    same=open(f'/tmp/dp/{name}_ref.wav','rb').read()==open(f'/tmp/dp/{name}_b.wav','rb').read()
    print(name,'identical',same); print('\n'.join(l for l in log.splitlines() if 'ixheaacd_b200' in l)[-1500:])
PY
