package com.particle_life.backend;

import java.io.*;
import java.util.*;

/**
 * Pins the C oracle (oracle/plife_oracle.c) against the REAL reference: runs the unmodified Physics.update() on the inputs of
 * the committed golden fixtures and writes what it produced, bit for bit.  Source only: there is no JVM in the build
 * environment.  With a JDK 21 and the reference's jars on the class path:
 *   python tests/golden/export_for_java.py                       (writes tests/golden/java/NAME.in.txt)
 *   java com.particle_life.backend.DumpVectors NAME.in.txt NAME.out.txt
 *   python tests/golden/check_java_vectors.py                    (compares NAME.out.txt with tests/golden/NAME.npz)
 * Input: "n m steps wrap rmax friction force dt", then m*m matrix entries, then n lines "x y vx vy type"; doubles as
 * Double.longBitsToDouble hex.  Output per step: n lines "x y vx vy type" in array order (the order update() leaves).
 * Only accelerator kind 0 (A/Main.java:275-280) exists in the reference; the lambda below is that one.
 */
public class DumpVectors {
    static double d(String hex) { return Double.longBitsToDouble(Long.parseUnsignedLong(hex, 16)); }
    static String h(double v) { return Long.toHexString(Double.doubleToRawLongBits(v)); }

    public static void main(String[] a) throws IOException {
        Scanner in = new Scanner(new File(a[0]));
        int n = in.nextInt(), m = in.nextInt(), steps = in.nextInt(), wrap = in.nextInt();
        Accelerator acc = (v, pos) -> {  // A/Main.java:275-280
            double beta = 0.3, dist = pos.length();
            double force = dist < beta ? (dist / beta - 1) : v * (1 - Math.abs(1 + beta - 2 * dist) / (1 - beta));
            return pos.mul(force / dist);
        };
        Physics p = new Physics(acc);
        p.settings.wrap = wrap != 0;
        p.settings.rmax = d(in.next());
        p.settings.friction = d(in.next());
        p.settings.force = d(in.next());
        p.settings.dt = d(in.next());
        p.preferredNumberOfThreads = 1;  // the result does not depend on it (each particle writes only itself)
        DefaultMatrix M = new DefaultMatrix(m);
        for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) M.set(i, j, d(in.next()));
        p.settings.matrix = M;
        Particle[] ps = new Particle[n];
        for (int i = 0; i < n; i++) {
            ps[i] = new Particle();
            ps[i].position.set(d(in.next()), d(in.next()), 0);
            ps[i].velocity.set(d(in.next()), d(in.next()), 0);
            ps[i].type = in.nextInt();
        }
        p.particles = ps;
        try (PrintWriter out = new PrintWriter(new BufferedWriter(new FileWriter(a[1])))) {
            for (int s = 0; s < steps; s++) {
                p.update();
                for (Particle q : p.particles)
                    out.println(h(q.position.x) + " " + h(q.position.y) + " " + h(q.velocity.x) + " " + h(q.velocity.y) + " " + q.type);
            }
        }
        p.kill();
    }
}
