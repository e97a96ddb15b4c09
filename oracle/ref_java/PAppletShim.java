/*
 * PAppletShim.java -- a headless stand-in for the part of processing.core.PApplet that the UNMODIFIED Lilypad tabs on
 * the AFCCylinder.update2() path use (AFCCylinder, BDIM, Body, BodyUnion, Field, MG, OrthoNormal, PoissonMatrix,
 * SaveScalar, VectorField, Window .pde).  TEST INFRASTRUCTURE: it exists so that somebody with a JDK can run the
 * reference's own Java arithmetic and produce reference-pinned golden files (see README.md in this directory); the
 * build container of this repository has no JVM, so this file has never been compiled here.
 *
 * Arithmetic helpers restate processing.core.PApplet / PVector (Processing 3.x), which are thin wrappers over
 * java.lang.Math evaluated in double and narrowed to float -- the only third-party arithmetic on the path (SURVEY 8c).
 * Drawing calls are no-ops.
 */
import java.io.BufferedReader;
import java.io.File;
import java.io.FileReader;
import java.io.FileWriter;
import java.io.IOException;
import java.io.PrintWriter;
import java.util.ArrayList;

public class PAppletShim {
  // ---- constants (processing.core.PConstants) ----
  public static final float PI = (float) Math.PI;
  public static final float TWO_PI = (float) (2.0 * Math.PI);
  public static final float HALF_PI = (float) (Math.PI / 2.0);
  public static final int RGB = 1, LEFT = 37, BASELINE = 0, CLOSE = 2;
  public int width = 512, height = 256;            // clientCFD.pde:19 size(512,256)
  public int mouseX = 0, mouseY = 0;
  public boolean quiet = true;                     // drop the per-step "t=.. drag=.. lift=.." prints of AFCCylinder.update2
  public String sketchPath = ".";                  // directory that holds saved/init/init.bdim

  // ---- arithmetic: PApplet.java ----
  public static float sin(float a) { return (float) Math.sin(a); }
  public static float cos(float a) { return (float) Math.cos(a); }
  public static float sqrt(float a) { return (float) Math.sqrt(a); }
  public static float atan2(float y, float x) { return (float) Math.atan2(y, x); }
  public static float pow(float a, float b) { return (float) Math.pow(a, b); }
  public static float sq(float a) { return a * a; }
  public static float abs(float a) { return (a < 0) ? -a : a; }
  public static int abs(int a) { return (a < 0) ? -a : a; }
  public static float min(float a, float b) { return (a < b) ? a : b; }
  public static float max(float a, float b) { return (a > b) ? a : b; }
  public static int min(int a, int b) { return (a < b) ? a : b; }
  public static int max(int a, int b) { return (a > b) ? a : b; }
  public static float min(float a, float b, float c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
  public static float max(float a, float b, float c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }
  public static float mag(float a, float b) { return (float) Math.sqrt(a * a + b * b); }
  public static int round(float a) { return Math.round(a); }
  public static float map(float v, float s1, float e1, float s2, float e2) { return s2 + (e2 - s2) * ((v - s1) / (e1 - s1)); }
  public static float constrain(float a, float lo, float hi) { return (a < lo) ? lo : ((a > hi) ? hi : a); }

  // ---- conversions: the Processing preprocessor turns float(x) / int(x) / str(x) into these ----
  public static float parseFloat(String s) { try { return Float.parseFloat(s); } catch (NumberFormatException e) { return Float.NaN; } }
  public static float parseFloat(int v) { return (float) v; }
  public static float parseFloat(float v) { return v; }
  public static float[] parseFloat(String[] s) { float[] r = new float[s.length]; for (int i = 0; i < s.length; i++) r[i] = parseFloat(s[i]); return r; }
  public static int parseInt(float v) { return (int) v; }
  public static int parseInt(String s) { try { return Integer.parseInt(s.trim()); } catch (NumberFormatException e) { return 0; } }
  public static String str(int v) { return String.valueOf(v); }
  public static String str(float v) { return String.valueOf(v); }
  public static String[] split(String s, char delim) {
    ArrayList<String> out = new ArrayList<String>();
    int start = 0;
    for (int i = 0; i < s.length(); i++) if (s.charAt(i) == delim) { out.add(s.substring(start, i)); start = i + 1; }
    out.add(s.substring(start));
    return out.toArray(new String[0]);
  }
  public static String nfs(float v, int left, int right) {       // sign-padded fixed format; printing only
    String body = String.format("%0" + (left + right + 1) + "." + right + "f", Math.abs(v));
    return ((v < 0) ? "-" : " ") + body;
  }

  // ---- console / lifecycle ----
  public void print(Object o) { if (!quiet) System.out.print(o); }
  public void println(Object o) { if (!quiet) System.out.println(o); }
  public void println() { if (!quiet) System.out.println(); }
  public void exit() { throw new RuntimeException("sketch called exit()"); }

  // ---- files: paths in the sketch use Windows separators ("saved\\init\\init.bdim", AFCCylinder.pde:37) ----
  File resolve(String name) {
    File f = new File(name.replace('\\', '/'));
    return f.isAbsolute() ? f : new File(sketchPath, f.getPath());
  }
  public PrintWriter createWriter(String name) {
    try {
      File f = resolve(name);
      if (f.getParentFile() != null) f.getParentFile().mkdirs();
      return new PrintWriter(new FileWriter(f));
    } catch (IOException e) { throw new RuntimeException(e); }
  }
  public String[] loadStrings(String name) {
    try (BufferedReader r = new BufferedReader(new FileReader(resolve(name)))) {
      ArrayList<String> out = new ArrayList<String>();
      for (String ln = r.readLine(); ln != null; ln = r.readLine()) out.add(ln);
      return out.toArray(new String[0]);
    } catch (IOException e) { throw new RuntimeException(e); }
  }

  // ---- drawing: no-ops ----
  public static class PFont { }
  public static class PImage { public int[] pixels; PImage(int w, int h) { pixels = new int[Math.max(w * h, 0)]; } public void loadPixels() { } public void updatePixels() { } }
  public static class MouseEvent { public int getCount() { return 0; } }
  public PFont loadFont(String name) { return new PFont(); }
  public PImage createImage(int w, int h, int mode) { return new PImage(w, h); }
  public int color(float gray) { int g = (int) constrain(gray, 0, 255); return 0xFF000000 | (g << 16) | (g << 8) | g; }
  public void colorMode(int mode, float max) { }
  public void background(float g) { }
  public void background(float r, float g, float b) { }
  public void fill(int c) { }
  public void stroke(int c) { }
  public void noStroke() { }
  public void strokeWeight(float w) { }
  public void beginShape() { }
  public void endShape(int mode) { }
  public void endShape() { }
  public void vertex(float x, float y) { }
  public void line(float a, float b, float c, float d) { }
  public void ellipse(float a, float b, float c, float d) { }
  public void rect(float a, float b, float c, float d) { }
  public void image(PImage img, float a, float b, float c, float d) { }
  public void textFont(PFont f) { }
  public void textAlign(int a, int b) { }
  public void text(String s, float x, float y) { }
  public void pushMatrix() { }
  public void popMatrix() { }
  public void translate(float x, float y) { }
  public void rotate(float a) { }
  public void smooth() { }

  // ---- processing.core.PVector (3.x: the mutators return this) ----
  public static class PVector {
    public float x, y, z;
    public PVector() { }
    public PVector(float x, float y) { this.x = x; this.y = y; }
    public PVector(float x, float y, float z) { this.x = x; this.y = y; this.z = z; }
    public PVector copy() { return new PVector(x, y, z); }
    public PVector get() { return copy(); }
    public PVector set(float x, float y) { this.x = x; this.y = y; return this; }
    public float mag() { return (float) Math.sqrt(x * x + y * y + z * z); }
    public PVector add(PVector v) { x += v.x; y += v.y; z += v.z; return this; }
    public PVector add(float a, float b) { x += a; y += b; return this; }
    public PVector add(float a, float b, float c) { x += a; y += b; z += c; return this; }
    public PVector sub(PVector v) { x -= v.x; y -= v.y; z -= v.z; return this; }
    public PVector mult(float n) { x *= n; y *= n; z *= n; return this; }
    public PVector div(float n) { x /= n; y /= n; z /= n; return this; }
    public static PVector add(PVector a, PVector b) { return new PVector(a.x + b.x, a.y + b.y, a.z + b.z); }
    public static PVector sub(PVector a, PVector b) { return new PVector(a.x - b.x, a.y - b.y, a.z - b.z); }
    public static PVector mult(PVector a, float n) { return new PVector(a.x * n, a.y * n, a.z * n); }
    public static PVector div(PVector a, float n) { return new PVector(a.x / n, a.y / n, a.z / n); }
  }
}
