/*
 * HeadlessLilypad.java -- drives the reference's own AFCCylinder (compiled from the unmodified .pde tabs by
 * build_ref.sh) without a display and without the XML-RPC agent, and writes the SAME golden files that
 * tests/golden/make_oracle_goldens.py writes from the C oracle:
 *
 *   config1_trace_1000.bin   BASELINE config 1: from saved/init/init.bdim, 1000 solver steps (AFCCylinder.update2), the
 *                            action of RL step k, a_k = (0.8 sin(2 pi k/25), -0.8 sin(2 pi k/25 + 1)) narrowed to float,
 *                            applied at solver step 16 k (xi_m = 5 xi, clientCFD.pde:53-54); per solver step the raw force
 *                            (fx, fy) and the 32 probes of SaveScalar.addData03 (SaveScalar.pde:61-72): little-endian
 *                            float32 [1000][34]
 *   config1_fields_100.bin   ux, uy, p (386 x 194 each, i-major, ghosts included) after 100 solver steps with the constant
 *                            action (0.5, -0.3): little-endian float32 [3][386][194]
 *   config1_obs_64.txt       the (Cl, Cd) observations of the first 64 RL steps with the clientCFD.draw() accumulation
 *                            (clientCFD.pde:39-55, initTime = -1 so that accumulation starts at the first step, as the
 *                            oracle CLI and the parity tests do), printed with Float.toString
 *
 * usage: java -cp oracle/_ref HeadlessLilypad <dir with saved/init/init.bdim> <output dir>
 * Compare with the committed goldens: python oracle/ref_java/compare_ref.py <output dir>
 *
 * TEST INFRASTRUCTURE.  Never compiled in this repository's build container (no JDK there).
 */
import java.io.DataOutputStream;
import java.io.File;
import java.io.FileOutputStream;
import java.io.IOException;
import java.io.PrintWriter;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;

public class HeadlessLilypad {
  static float[] action(int k) {
    return new float[] {(float) (0.8 * Math.sin(2 * Math.PI * k / 25.0)), (float) (-0.8 * Math.sin(2 * Math.PI * k / 25.0 + 1.0))};
  }

  static LilypadSketch.AFCCylinder make(LilypadSketch s) {
    // clientCFD.pde:93-102 setUpNewSim: resolution 24, Re 500, dR .125, gR .2, theta PI/3, xi 0, tStep .0075, 16 x 8, resume
    int resolution = 24, zoom = 100 / resolution;
    return s.new AFCCylinder(resolution, 500, .125f, .2f, PAppletShim.PI / 3, 0f, 0f, .0075f, 16, 8, zoom, true);
  }

  static void setAction(LilypadSketch.AFCCylinder t, float a1, float a2) {    // clientCFD.pde:51-54
    t.xi1 = a1; t.xi2 = a2; t.xi1_m = 5 * t.xi1; t.xi2_m = 5 * t.xi2;
  }

  static void putFloats(DataOutputStream out, float[] v) throws IOException {
    ByteBuffer b = ByteBuffer.allocate(4 * v.length).order(ByteOrder.LITTLE_ENDIAN);
    for (float f : v) b.putFloat(f);
    out.write(b.array());
  }

  public static void main(String[] args) throws IOException {
    LilypadSketch s = new LilypadSketch();
    s.sketchPath = args.length > 0 ? args[0] : ".";
    File outDir = new File(args.length > 1 ? args[1] : ".");
    outDir.mkdirs();

    // ---- trace: 1000 solver steps, action k at solver step 16 k ----
    LilypadSketch.AFCCylinder t = make(s);
    float D = 24, centX = (16 * 24) / 4f, centY = (8 * 24) / 2f;             // SaveScalar.pde:35-41
    try (DataOutputStream out = new DataOutputStream(new FileOutputStream(new File(outDir, "config1_trace_1000.bin")))) {
      float[] row = new float[34];
      for (int step = 0; step < 1000; step++) {
        if (step % 16 == 0) { float[] a = action(step / 16); setAction(t, a[0], a[1]); }
        t.update2();
        row[0] = t.force.x; row[1] = t.force.y;
        for (int i = 0; i < 32; i++) {                                       // SaveScalar.pde:64-67
          float xPre = PAppletShim.cos((float) i / 32 * PAppletShim.PI * 2) * D / 2 + centX;
          float yPre = PAppletShim.sin((float) i / 32 * PAppletShim.PI * 2) * D / 2 + centY;
          row[2 + i] = t.flow.p.linear(xPre, yPre);
        }
        putFloats(out, row);
      }
    }

    // ---- fields after 100 steps with a constant action ----
    t = make(s);
    setAction(t, 0.5f, -0.3f);
    for (int step = 0; step < 100; step++) t.update2();
    try (DataOutputStream out = new DataOutputStream(new FileOutputStream(new File(outDir, "config1_fields_100.bin")))) {
      float[][][] f = {t.flow.u.x.a, t.flow.u.y.a, t.flow.p.a};
      for (float[][] a : f) for (float[] r : a) putFloats(out, r);
    }

    // ---- observations with the draw() accumulation (clientCFD.pde:39-55; Cd, Cl deliberately not zeroed) ----
    t = make(s);
    int callLearn = 16, k = 0;                                               // xi = 0 until the first observation, then a_0, a_1, ...
    float Cd = 0, Cl = 0;
    try (PrintWriter out = new PrintWriter(new File(outDir, "config1_obs_64.txt"))) {
      int produced = 0;
      while (produced < 64) {
        t.update2();
        callLearn--;
        Cd += t.force.x;
        Cl += t.force.y;
        if (callLearn <= 0) {
          callLearn = 16;
          Cd = Cd / callLearn * 2 / 24;
          Cl = Cl / callLearn * 2 / 24;
          out.println(Float.toString(Cl) + " " + Float.toString(Cd));
          produced++;
          float[] a = action(k++);
          setAction(t, a[0], a[1]);
        }
      }
    }
    System.out.println("written config1_trace_1000.bin, config1_fields_100.bin, config1_obs_64.txt to " + outDir);
  }
}
