// ReplayHarness.cs — the executed-reference pin of the parity claim (DESIGN.md section 7), for the day a .NET box exists.
// NOT COMPILED IN THIS REPOSITORY (no .NET SDK in the image).  A SLAM.NET maintainer
//   1. applies dotnet/reference_verification_hook.patch to CoreSLAM/CoreSLAMProcessor.cs (adds CandidateTable, LastDistance),
//   2. adds this file, CoreSLAMProcessor.B200.cs and CoreSlamNative.cs as a console project referencing CoreSLAM.csproj,
//   3. runs   ReplayHarness <log.cslg> [--golden out.golden] [--b200]
// The harness replays one CSLG scan log (include/coreslam_b200.h, "Scan-log files"; parameters in <log.cslg>.json, written
// by tools/make_cslg_fixture.py) through the REFERENCE CoreSLAMProcessor with the log's candidate tables, and
//   --golden   writes one line per scan: scan index, the pose as three float bit patterns (hex), the winning distance, and
//              the CRC-32 of the little-endian HoleMap — the golden file tests/test_golden_reference.py checks the CPU
//              oracle and the CUDA path against (tests/golden/<name>.cslg.golden);
//   --b200     also runs the B200 drop-in scan by scan and stops at the first divergence (pose bits, distance, map CRC).
//
// The log stores clouds (ScanCloud.Points) and odometry poses.  The reference only accepts List<ScanSegment>, so each scan
// is fed as ONE segment at the odometry pose with rays (atan2(y, x), |p|): ScanSegmentsToCloud (:187-207) then rebuilds the
// cloud from polar form — in the reference, in the drop-in, and in the pytest that consumes the golden file alike.
using System;
using System.Collections.Generic;
using System.Globalization;
using System.IO;
using System.Numerics;
using System.Text.Json;
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

        static uint Crc32(ushort[] pixels)   // zlib polynomial, over the little-endian bytes (KAT-B of SURVEY 8c uses the same)
        {
            uint crc = 0xFFFFFFFFu;
            foreach (ushort v in pixels)
                for (int b = 0; b < 2; b++)
                {
                    crc ^= (uint)((v >> (8 * b)) & 0xFF);
                    for (int i = 0; i < 8; i++) crc = (crc >> 1) ^ (0xEDB88320u & (uint)-(int)(crc & 1));
                }
            return ~crc;
        }

        static string Bits(float f) => BitConverter.SingleToInt32Bits(f).ToString("x8", CultureInfo.InvariantCulture);

        public static int Main(string[] args)
        {
            string logPath = args[0];
            string goldenPath = null;
            bool runB200 = false;
            for (int i = 1; i < args.Length; i++)
            {
                if (args[i] == "--golden") goldenPath = args[++i];
                else if (args[i] == "--b200") runB200 = true;
            }
            List<Scan> scans = ReadLog(logPath, out int maxPoints, out int nOffsets);
            using JsonDocument cfg = JsonDocument.Parse(File.ReadAllText(logPath + ".json"));
            JsonElement c = cfg.RootElement;
            float phys = c.GetProperty("physical_map_size").GetSingle();
            int holeSize = c.GetProperty("hole_map_size").GetInt32(), obstSize = c.GetProperty("obstacle_map_size").GetInt32();
            int iters = c.GetProperty("iterations_per_thread").GetInt32(), threads = c.GetProperty("num_search_threads").GetInt32();
            float sxy = c.GetProperty("sigma_xy").GetSingle(), sth = c.GetProperty("sigma_theta").GetSingle();
            if (nOffsets != iters * Math.Max(threads, 1)) throw new InvalidDataException("log offsets != T*I of its .json");

            using var reference = new CoreSLAM.CoreSLAMProcessor(phys, holeSize, obstSize, scans[0].Odometry, sxy, sth, iters, threads);
            CoreSLAMProcessor b200 = runB200
                ? new CoreSLAMProcessor(phys, holeSize, obstSize, scans[0].Odometry, sxy, sth, iters, threads, maxPoints: maxPoints) { SyncMapAfterUpdate = true }
                : null;
            using StreamWriter golden = goldenPath != null ? new StreamWriter(goldenPath) : null;
            golden?.WriteLine("# scan pose_x pose_y pose_theta (float bits, hex) distance holemap_crc32   -- reference mikkleini/slam.net, CandidateTable hook");
            for (int k = 0; k < scans.Count; k++)
            {
                List<ScanSegment> segs = AsSegments(scans[k]);
                reference.CandidateTable = scans[k].Offsets;   // the verification hook (reference_verification_hook.patch)
                reference.Update(segs);
                uint crc = Crc32(reference.HoleMap.Pixels);
                golden?.WriteLine($"{k} {Bits(reference.Pose.X)} {Bits(reference.Pose.Y)} {Bits(reference.Pose.Z)} {reference.LastDistance} {crc:x8}");
                if (b200 == null) continue;
                b200.Update(segs, scans[k].Offsets);
                bool same = Bits(reference.Pose.X) == Bits(b200.Pose.X) && Bits(reference.Pose.Y) == Bits(b200.Pose.Y) &&
                            Bits(reference.Pose.Z) == Bits(b200.Pose.Z) && reference.LastDistance == b200.LastDistance &&
                            crc == Crc32(b200.HoleMap.Pixels);
                if (!same)
                {
                    Console.WriteLine($"scan {k}: reference pose {reference.Pose} distance {reference.LastDistance} crc {crc:x8}; " +
                                      $"b200 pose {b200.Pose} distance {b200.LastDistance} crc {Crc32(b200.HoleMap.Pixels):x8}");
                    return 1;
                }
            }
            b200?.Dispose();
            Console.WriteLine($"{scans.Count} scans replayed" + (runB200 ? ": poses, distances and HoleMap identical" : "") +
                              (goldenPath != null ? $"; golden written to {goldenPath}" : ""));
            return 0;
        }
    }
}
