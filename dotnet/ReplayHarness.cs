// ReplayHarness.cs — replays one CSLG scan log (include/coreslam_b200.h, "Scan-log files") through the reference
// CoreSLAMProcessor and through the B200 drop-in, scan by scan, and reports the first divergence.
// NOT COMPILED IN THIS REPOSITORY (no .NET SDK in the image).  A SLAM.NET maintainer adds it as a console project that
// references CoreSLAM.csproj; it is the missing executed-reference pin of the parity claim (DESIGN.md section 7).
//
// The log stores clouds (ScanCloud.Points) and odometry poses.  The reference only accepts List<ScanSegment>, so each scan
// is fed as ONE segment at the odometry pose with rays (atan2(y, x), |p|): ScanSegmentsToCloud (:187-207) then rebuilds
// the cloud from polar form in both implementations alike.  Candidate tables: the reference draws its own (unseeded)
// deviates, so bit-exact comparison needs the verification hook — replace the two sampler calls in MonteCarloSearch
// (:633-638) by reads from the table of the current scan (3 floats per candidate, thread t iteration i at 1 + t*I + i - 1).
using System;
using System.Collections.Generic;
using System.IO;
using System.Numerics;
using BaseSLAM;

namespace CoreSLAM.B200
{
    public static class ReplayHarness
    {
        public sealed class Scan
        {
            public Vector3 Odometry;
            public Vector2[] Points;
            public float[] Offsets;   // nOffsets * 3, or empty
        }

        public static List<Scan> ReadLog(string path, out int maxPoints, out int nOffsets)
        {
            using var r = new BinaryReader(File.OpenRead(path));
            if (new string(r.ReadChars(4)) != "CSLG" || r.ReadUInt32() != 1) throw new InvalidDataException("not a CSLG version 1 file");
            int nScans = (int)r.ReadUInt32();
            maxPoints = (int)r.ReadUInt32();
            nOffsets = (int)r.ReadUInt32();
            r.ReadBytes(12);
            var scans = new List<Scan>(nScans);
            for (int k = 0; k < nScans; k++)
            {
                int n = (int)r.ReadUInt32();
                var s = new Scan { Odometry = new Vector3(r.ReadSingle(), r.ReadSingle(), r.ReadSingle()), Points = new Vector2[n] };
                for (int i = 0; i < n; i++) s.Points[i] = new Vector2(r.ReadSingle(), r.ReadSingle());
                s.Offsets = new float[nOffsets * 3];
                for (int i = 0; i < s.Offsets.Length; i++) s.Offsets[i] = r.ReadSingle();
                scans.Add(s);
            }
            return scans;
        }

        static List<ScanSegment> AsSegments(Scan s)
        {
            var seg = new ScanSegment { Pose = s.Odometry, IsLast = true };
            foreach (Vector2 p in s.Points) seg.Rays.Add(new Ray(MathF.Atan2(p.Y, p.X), p.Length()));
            return new List<ScanSegment> { seg };
        }

        public static int Main(string[] args)
        {
            List<Scan> scans = ReadLog(args[0], out int maxPoints, out int nOffsets);
            float phys = 40.0f; int holeSize = 2048, obstSize = 512, iters = 1024, threads = 4;   // cfg2 of BASELINE.json
            using var reference = new CoreSLAM.CoreSLAMProcessor(phys, holeSize, obstSize, scans[0].Odometry, 0.1f, 0.17453292f, iters, threads);
            using var b200 = new CoreSLAMProcessor(phys, holeSize, obstSize, scans[0].Odometry, 0.1f, 0.17453292f, iters, threads, maxPoints: maxPoints);
            for (int k = 0; k < scans.Count; k++)
            {
                List<ScanSegment> segs = AsSegments(scans[k]);
                // reference.CandidateTable = scans[k].Offsets;   // the verification hook described above
                reference.Update(segs);
                b200.Update(segs /* , scans[k].Offsets through cs_update_segments' cand_offsets */);
                if (reference.Pose != b200.Pose)
                {
                    Console.WriteLine($"scan {k}: reference {reference.Pose} b200 {b200.Pose}");
                    return 1;
                }
            }
            b200.HoleMap.SyncToHost();
            for (int i = 0; i < b200.HoleMap.Pixels.Length; i++)
                if (b200.HoleMap.Pixels[i] != reference.HoleMap.Pixels[i]) { Console.WriteLine($"HoleMap cell {i} differs"); return 2; }
            Console.WriteLine($"{scans.Count} scans: poses and HoleMap identical");
            return 0;
        }
    }
}
