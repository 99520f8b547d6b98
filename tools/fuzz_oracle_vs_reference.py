#!/usr/bin/env python3
"""Randomised comparison of the CPU oracle with the compiled reference (oracle/_ref/*.so) -- a development tool, not part of the test
suite (needs /root/reference or a prebuilt oracle/_ref).  python tools/fuzz_oracle_vs_reference.py [match|tsdf|orb|linematch] [seconds]
Round 1, 150 s each on 8 cores: match 5534 iterations x 5 searches, tsdf 246 maps (~800 integrations), orb 1159 images: 0 mismatches.
Round 2: linematch 960 query / train sets (39 399 rows whose two neighbours are equidistant) in 9 s: 0 mismatches."""
import sys
which = sys.argv[1] if len(sys.argv) > 1 else "match"
SECONDS = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
import pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))

if which == "match":
    import sys, time, numpy as np
    from plvs_b200 import synth, scenario
    from plvs_b200.matcher import featvec
    from oracle import match as OM, orb as O
    K = synth.intrinsics(640, 480); tab = O.Tables(1500)
    frames={}
    def frame(f):
        if f not in frames:
            kp, desc, mono, _ = O.extract_port(synth.gray_frame(f), 1500)
            x = scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale); x.level_sigma2 = tab.sigma2; frames[f]=x
        return frames[f]
    t0=time.time(); it=0; bad=0
    rng=np.random.default_rng(123)
    while time.time()-t0 < SECONDS:
        a=int(rng.integers(0,30)); b=a+int(rng.integers(1,4))
        last,cur=frame(a),frame(b)
        seed=int(rng.integers(0,1<<30))
        qm,_=scenario.map_queries(last,cur,K,synth.pose(a),synth.pose(b),seed=seed)
        ql,_=scenario.last_queries(last,cur,K,synth.pose(a),synth.pose(b))
        r2=np.random.default_rng(seed)
        qm["flags"]=(r2.random(len(qm))<0.85).astype(np.uint32); ql["flags"]=(r2.random(len(ql))<0.85).astype(np.uint32)
        dup=int(r2.integers(1,4))
        qm=np.concatenate([qm]+[qm[::k] for k in range(2,2+dup)]); ql=np.concatenate([ql]+[ql[::k] for k in range(2,2+dup)])
        ql["invz"][::int(r2.integers(5,40))]*=-1
        claimed=(r2.random(cur.n)<r2.uniform(0,0.5)).astype(np.uint8)
        th=float(r2.choice([1.0,2.0,3.0,5.0,7.0,15.0,25.0]))
        ratio=float(r2.choice([0.6,0.8,0.9]))
        far=bool(r2.integers(0,2)); thf=float(r2.uniform(1,5))
        n,a1=OM.search_by_projection_map(cur,qm,th,ratio,far,thf,claimed); rn,ra=OM.ref_search_by_projection_map(cur,qm,th,ratio,far,thf,claimed)
        if n!=rn or not np.array_equal(a1,ra): bad+=1; print("MAP mismatch",a,b,seed,th)
        qc,z=OM.canonical_last_queries(ql)
        fwd,bwd=[(False,False),(True,False),(False,True)][int(r2.integers(0,3))]; chk=bool(r2.integers(0,2))
        n,a1=OM.search_by_projection_last(cur,qc,th,fwd,bwd,chk,claimed); rn,ra=OM.ref_search_by_projection_last(cur,qc,z,th,fwd,bwd,chk,claimed)
        if n!=rn or not np.array_equal(a1,ra): bad+=1; print("LAST mismatch",a,b,seed,th)
        nodes=int(r2.choice([8,64,256,1024]))
        fv1,fv2=featvec(scenario.node_ids(last.desc,nodes)),featvec(scenario.node_ids(cur.desc,nodes))
        h1=(r2.random(last.n)<0.5).astype(np.uint8); h2=(r2.random(cur.n)<0.5).astype(np.uint8)
        F12,ep=scenario.fundamental(K,synth.pose(a),synth.pose(b))
        co=bool(r2.integers(0,2)); os_=bool(r2.integers(0,2))
        n,m=OM.search_for_triangulation(last,cur,fv1,fv2,h1,h2,F12,ep,os_,co,chk); rn,rm=OM.ref_search_for_triangulation(last,cur,fv1,fv2,h1,h2,F12,ep,os_,co,chk)
        if n!=rn or not np.array_equal(m,rm): bad+=1; print("TRI mismatch",a,b,seed)
        n,m=OM.search_by_bow(last,cur,fv1,fv2,h1,ratio,chk); rn,rm=OM.ref_search_by_bow(last,cur,fv1,fv2,h1,ratio,chk)
        if n!=rn or not np.array_equal(m,rm): bad+=1; print("BOW mismatch",a,b,seed)
        n,m=OM.search_by_bow_kf(last,cur,fv1,fv2,h1,h2,ratio,chk); rn,rm=OM.ref_search_by_bow_kf(last,cur,fv1,fv2,h1,h2,ratio,chk)
        if n!=rn or not np.array_equal(m,rm): bad+=1; print("BOWKF mismatch",a,b,seed)
        it+=1
    print("iterations",it,"mismatches",bad)
if which == "tsdf":
    import sys, os, time, numpy as np
    from plvs_b200 import synth, scenario, tsdf as T
    from oracle import tsdf as OT
    devnull=os.open(os.devnull,os.O_WRONLY); saved=os.dup(1)
    def quiet(on):
        sys.stdout.flush()
        os.dup2(devnull if on else saved,1)
    rng=np.random.default_rng(7)
    t0=time.time(); it=0; bad=0
    w,h=96,72
    K=synth.intrinsics(w,h)
    while time.time()-t0< SECONDS:
        kw=dict(voxel_resolution=float(rng.choice([0.04,0.05,0.08])), near_plane=0.1, far_plane=float(rng.uniform(2.5,4.5)), use_color=int(rng.integers(0,2)), use_carving=int(rng.integers(0,2)),
                carving_dist=float(rng.uniform(0,0.1)), trunc_scale=float(rng.uniform(3,8)))
        p=T.default_params(max_blocks=8192,**kw)
        o=OT.Map(p,threads=8); r=OT.RefMap(p)
        for m in (o,r): m.set_camera(K["fx"],K["fy"],K["cx"],K["cy"],w,h)
        for s in range(int(rng.integers(2,5))):
            f=int(rng.integers(0,40))
            d=synth.depth_frame(f,w,h).copy(); c=synth.bgr_frame(f,w,h)
            d+=rng.normal(0,0.01,d.shape).astype(np.float32)*(rng.random()<0.5)
            d[rng.random(d.shape)<0.03]=np.nan; d[rng.random(d.shape)<0.03]=0
            if rng.random()<0.3: d=np.maximum(d-np.float32(rng.uniform(0.1,0.6)),0).astype(np.float32)
            P=synth.pose(f).copy().reshape(3,4)
            a=rng.normal(0,0.4,3); ang=np.linalg.norm(a)
            if ang>1e-6:
                k=a/ang; Kx=np.array([[0,-k[2],k[1]],[k[2],0,-k[0]],[-k[1],k[0],0]]); R=np.eye(3)+np.sin(ang)*Kx+(1-np.cos(ang))*Kx@Kx
                P[:,:3]=(P[:,:3].astype(np.float64)@R).astype(np.float32)
            P[:,3]+=rng.normal(0,0.2,3).astype(np.float32)
            route=rng.random()
            quiet(True)
            if route<0.6 or not kw["use_color"]:
                col=c if kw["use_color"] else None
                o.integrate(d,P,col); r.integrate(d,P,col)
            else:
                dd=np.nan_to_num(d,nan=0.0).astype(np.float32)
                xyz,rgb=scenario.cloud_from_depth(dd,c,K,step=int(rng.integers(1,3)))
                o.integrate_cloud(xyz,rgb,P,dd); r.integrate_cloud(xyz,rgb,P,dd)
            quiet(False)
            ok_,os_,ow,oc=o.download(); rk,rs,rw,rc=r.download()
            same=np.array_equal(ok_,rk) and np.array_equal(os_,rs) and np.array_equal(ow,rw) and np.array_equal(oc,rc)
            if not same:
                bad+=1; print("MISMATCH",kw,f,route, len(ok_),len(rk)); break
        it+=1
    print("maps",it,"mismatches",bad)
if which == "orb":
    import sys, time, numpy as np
    from plvs_b200 import synth
    from oracle import orb as O
    import cv2
    rng=np.random.default_rng(99)
    t0=time.time(); it=0; bad=0
    while time.time()-t0< SECONDS:
        w=int(rng.integers(160,700)); h=int(rng.integers(120,500)); nf=int(rng.integers(100,2500))
        f=int(rng.integers(0,60))
        img=synth.gray_frame(f,w,h)
        mode=rng.integers(0,4)
        if mode==1: img=cv2.GaussianBlur(img,(0,0),float(rng.uniform(0.5,3)))
        elif mode==2: img=np.clip(img.astype(np.int16)+rng.integers(-30,30,img.shape),0,255).astype(np.uint8)
        elif mode==3: img=rng.integers(0,256,(h,w),dtype=np.uint8)
        lap=(0,0) if rng.random()<0.7 else (int(rng.integers(0,w//2)),int(rng.integers(w//2,w)))
        r=O.RefExtractor(nf)(img,lap); o=O.extract_port(img,nf,lapping=lap)
        ok=len(r[0])==len(o[0]) and all(np.array_equal(r[0][k],o[0][k]) for k in ("x","y","size","angle","response","octave")) and np.array_equal(r[1],o[1]) and r[2]==o[2]
        if not ok: bad+=1; print("MISMATCH",w,h,nf,f,mode,lap)
        it+=1
    print("images",it,"mismatches",bad)

if which == "linematch":
    import time, numpy as np
    from oracle import linematch as L
    from tests.linematch_cases import cases
    R = L.RefLineMatcher()
    t0 = time.time(); bad = 0; n = 0; ties = 0; seed = 1000
    while time.time() - t0 < SECONDS:
        for q, t, mask in cases(24, seed=seed, nq_max=200, nt_max=300):
            a = L.knn2(q, t, mask, 0.78); b = R.knn2(q, t, mask, 0.78)
            bad += not (all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4])) and a[4] == b[4]); n += 1
            ties += int(np.sum(a[2][:, 0] == a[2][:, 1]))
        seed += 1
    print(f"linematch: {n} query / train sets, {ties} rows with equidistant neighbours, {bad} mismatches")
