package com.particle_life.backend;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import org.joml.Vector3d;

/**
 * Drop-in replacement for {@link Physics} that runs {@code update()} on a B200 through libplife.so.
 *
 * <p>UNBUILT / UNTESTED in the build environment (no JVM there; SURVEY.md F7/F8).  It targets the
 * Java FFM API (final in JDK 22; on the app's JDK 21 toolchain compile with {@code --enable-preview}
 * or use the JNI variant sketched in INTEGRATION.md).  It binds exactly the {@code plife_*}
 * symbols of include/plife.h.
 *
 * <p>Usage in {@code Main.createPhysics()} (A/Main.java:281-285): construct this class instead of
 * {@code ExtendedPhysics}.  {@code Loop} and everything above stay unchanged: they call
 * {@code update()}, mutate {@code settings}, and read {@code particles} when a snapshot is due.
 * The particle buffers live on the GPU; {@link #requestSnapshot()} / {@link #finishSnapshot()} hand the renderer
 * float buffers without stalling the physics (A/PhysicsSnapshot.java:24-62), {@link #pull()} refreshes the Java-side
 * {@code particles} array when Java code wants to edit it, and {@link #push()} uploads the edits.
 */
public class NativePhysics extends Physics {

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup("libplife.so", Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    private static final ValueLayout.OfInt I32 = ValueLayout.JAVA_INT;
    private static final ValueLayout.OfLong I64 = ValueLayout.JAVA_LONG;
    private static final ValueLayout.OfDouble F64 = ValueLayout.JAVA_DOUBLE;
    private static final AddressLayout PTR = ValueLayout.ADDRESS;

    private static final MethodHandle CREATE = fn("plife_create", FunctionDescriptor.of(I32, PTR, PTR));
    private static final MethodHandle DESTROY = fn("plife_destroy", FunctionDescriptor.of(I32, PTR));
    private static final MethodHandle SET_SETTINGS = fn("plife_set_settings", FunctionDescriptor.of(I32, PTR, PTR));
    private static final MethodHandle SET_MATRIX = fn("plife_set_matrix", FunctionDescriptor.of(I32, PTR, I32, PTR));
    private static final MethodHandle SET_ACCELERATOR = fn("plife_set_accelerator", FunctionDescriptor.of(I32, PTR, I32, PTR, I32));
    private static final MethodHandle UPLOAD = fn("plife_upload", FunctionDescriptor.of(I32, PTR, I64, PTR, PTR, PTR, PTR));
    private static final MethodHandle DOWNLOAD = fn("plife_download", FunctionDescriptor.of(I32, PTR, PTR, PTR, PTR, PTR));
    private static final MethodHandle STEP = fn("plife_step", FunctionDescriptor.of(I32, PTR, F64, I32));
    private static final MethodHandle SYNC = fn("plife_sync", FunctionDescriptor.of(I32, PTR));
    private static final MethodHandle REQUEST_STOP = fn("plife_request_stop", FunctionDescriptor.of(I32, PTR));
    private static final MethodHandle LAST_ERROR = fn("plife_last_error", FunctionDescriptor.of(PTR, PTR));
    private static final MethodHandle COUNT = fn("plife_count", FunctionDescriptor.of(I64, PTR));
    private static final MethodHandle TYPE_HISTOGRAM = fn("plife_type_histogram", FunctionDescriptor.of(I32, PTR, PTR));
    private static final MethodHandle SNAPSHOT_ASYNC_U8 = fn("plife_snapshot_async_u8", FunctionDescriptor.of(I32, PTR, PTR, PTR, PTR));
    private static final MethodHandle SNAPSHOT_WAIT = fn("plife_snapshot_wait", FunctionDescriptor.of(I32, PTR));

    /** struct plife_config { int32 device, precision; int64 capacity; int32 flags, bins; void* stream; } */
    private static final MemoryLayout CONFIG = MemoryLayout.structLayout(I32.withName("device"), I32.withName("precision"),
            I64.withName("capacity"), I32.withName("flags"), I32.withName("bins"), PTR.withName("stream"));
    /** struct plife_settings { double rmax, friction, force; int32 wrap, reserved; } */
    private static final MemoryLayout SETTINGS = MemoryLayout.structLayout(F64.withName("rmax"), F64.withName("friction"),
            F64.withName("force"), I32.withName("wrap"), I32.withName("reserved"));

    private final Arena arena = Arena.ofShared(); // lives as long as the handle: settings block, matrix
    private final MemorySegment settingsBlock = arena.allocate(SETTINGS);
    private MemorySegment matrixBlock = MemorySegment.NULL; // re-allocated when the matrix size changes
    private int matrixBlockSize = -1;
    private MemorySegment handle;
    private boolean dirty = true; // Java-side particles changed since the last upload

    /** Float snapshot for the renderer (A/PhysicsSnapshot.java + A/ParticleRenderer.java): xy, vxy as float, type as byte. */
    public static final class Snapshot {
        public MemorySegment positions, velocities, types; // n*2 floats, n*2 floats, n bytes; off-heap, GL-uploadable
        public int particleCount;
        Arena arena; // owns the three buffers: closed and replaced when the particle count changes (brush / delete every frame)
    }

    private final Snapshot[] snapshots = {new Snapshot(), new Snapshot()};
    private int snapshotIndex = 0;

    public NativePhysics(Accelerator accelerator, PositionSetter positionSetter, MatrixGenerator matrixGenerator,
                         TypeSetter typeSetter) {
        super(accelerator, positionSetter, matrixGenerator, typeSetter); // generates matrix + 10000 particles on the host
        try (Arena tmp = Arena.ofConfined()) {
            MemorySegment cfg = tmp.allocate(CONFIG); // zero-initialised: device 0, PLIFE_F32, library-owned stream
            cfg.set(I64, 8, particles.length);
            MemorySegment out = tmp.allocate(PTR);
            check((int) CREATE.invoke(cfg, out));
            handle = out.get(PTR, 0);
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** Physics.update() (B/Physics.java:112): settings -> device, one step on the GPU. */
    @Override
    public void update() {
        try {
            // Order matters when the matrix SHRANK (setMatrixSize + ensureTypes ran on the Java side): the resident particles
            // still hold types >= m, so plife_set_matrix would refuse (PLIFE_ERR_STATE).  Upload the retyped particles first -
            // they validate against the old, larger matrix - then the matrix.  When it grew: matrix first, so that the new
            // types validate.
            boolean shrank = settings.matrix.size() < matrixBlockSize;
            if (dirty && shrank) push();
            pushSettings();
            if (dirty) push();
            check((int) STEP.invoke(handle, settings.dt, 1));
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    private void pushSettings() throws Throwable {
        MemorySegment s = settingsBlock;
        s.set(F64, 0, settings.rmax);
        s.set(F64, 8, settings.friction);
        s.set(F64, 16, settings.force);
        s.set(I32, 24, settings.wrap ? 1 : 0);
        check((int) SET_SETTINGS.invoke(handle, s));
        int m = settings.matrix.size();
        if (m != matrixBlockSize) { // a handful of sizes over the life of the app: the shared arena keeps them
            matrixBlock = arena.allocate(F64, (long) m * m);
            matrixBlockSize = m;
        }
        MemorySegment mat = matrixBlock;
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++) mat.setAtIndex(F64, (long) i * m + j, settings.matrix.get(i, j));
        check((int) SET_MATRIX.invoke(handle, m, mat));
        // a Java lambda cannot run on the device: kind 0 is the accelerator of A/Main.java:275-280
        check((int) SET_ACCELERATOR.invoke(handle, 0, MemorySegment.NULL, 0));
    }

    /** Upload the Java-side particle array (after setParticleCount / cursor edits / load). */
    public void push() throws Throwable {
        int n = particles.length;
        try (Arena tmp = Arena.ofConfined()) {
        MemorySegment pos = tmp.allocate(F64, 2L * n), vel = tmp.allocate(F64, 2L * n), typ = tmp.allocate(I32, n);
        for (int i = 0; i < n; i++) {
            Particle p = particles[i];
            pos.setAtIndex(F64, 2L * i, p.position.x);
            pos.setAtIndex(F64, 2L * i + 1, p.position.y);
            vel.setAtIndex(F64, 2L * i, p.velocity.x);
            vel.setAtIndex(F64, 2L * i + 1, p.velocity.y);
            typ.setAtIndex(I32, i, p.type);
        }
        check((int) UPLOAD.invoke(handle, (long) n, pos, vel, typ, MemorySegment.NULL));
        }
        dirty = false;
    }

    /** Display-time handoff: refresh the Java-side particles from the GPU (cell-sorted order, like the reference). */
    public void pull() throws Throwable {
        int n = particles.length;
        try (Arena tmp = Arena.ofConfined()) {
            MemorySegment pos = tmp.allocate(F64, 2L * n), vel = tmp.allocate(F64, 2L * n), typ = tmp.allocate(I32, n);
            check((int) DOWNLOAD.invoke(handle, pos, vel, typ, MemorySegment.NULL));
            for (int i = 0; i < n; i++) {
                Particle p = particles[i];
                p.position.set(pos.getAtIndex(F64, 2L * i), pos.getAtIndex(F64, 2L * i + 1), 0);
                p.velocity.set(vel.getAtIndex(F64, 2L * i), vel.getAtIndex(F64, 2L * i + 1), 0);
                p.type = typ.getAtIndex(I32, i);
            }
        }
    }

    /**
     * What Main.java:600-603 does with {@code loop.doOnce(() -> physicsSnapshot.take(...))}: start a float snapshot of the
     * current state.  The copy overlaps the following {@code update()} calls; {@link #finishSnapshot()} returns it.
     * Two buffer sets alternate, so the renderer may keep reading the previous snapshot meanwhile.
     */
    public void requestSnapshot() throws Throwable {
        int n = (int) (long) COUNT.invoke(handle);
        Snapshot s = snapshots[snapshotIndex];
        if (s.particleCount != n || s.positions == null) { // size changed: free this set's buffers, allocate new ones
            if (s.arena != null) s.arena.close(); // (its previous snapshot was handed out two requests ago: the renderer is done with it)
            s.arena = Arena.ofShared();
            s.positions = s.arena.allocate(ValueLayout.JAVA_FLOAT, 2L * n);
            s.velocities = s.arena.allocate(ValueLayout.JAVA_FLOAT, 2L * n);
            s.types = s.arena.allocate(ValueLayout.JAVA_BYTE, n);
            s.particleCount = n;
        }
        check((int) SNAPSHOT_ASYNC_U8.invoke(handle, s.positions, s.velocities, s.types));
    }

    public Snapshot finishSnapshot() throws Throwable {
        check((int) SNAPSHOT_WAIT.invoke(handle));
        Snapshot s = snapshots[snapshotIndex];
        snapshotIndex ^= 1;
        return s;
    }

    /** ExtendedPhysics.getTypeCount (A/ExtendedPhysics.java:19-26) without a download. */
    public int[] getTypeCount() throws Throwable {
        int m = settings.matrix.size();
        try (Arena tmp = Arena.ofConfined()) {
            MemorySegment out = tmp.allocate(I64, m);
            check((int) TYPE_HISTOGRAM.invoke(handle, out));
            int[] counts = new int[m];
            for (int i = 0; i < m; i++) counts[i] = (int) out.getAtIndex(I64, i);
            return counts;
        }
    }

    /** Call after any Java-side mutation of {@code particles} (A/Main.java:536-582, :671, :1320). */
    public void markDirty() {
        dirty = true;
    }

    @Override
    public void forceUpdateStop() {
        try {
            REQUEST_STOP.invoke(handle);
        } catch (Throwable ignored) {
        }
    }

    @Override
    public void kill() {
        try {
            if (handle != null) DESTROY.invoke(handle);
            handle = null;
            for (Snapshot s : snapshots) if (s.arena != null) s.arena.close();
            arena.close(); // settings block, matrix block
        } catch (Throwable ignored) {
        }
        super.kill();
    }

    private void check(int status) throws Throwable {
        if (status != 0) {
            MemorySegment msg = ((MemorySegment) LAST_ERROR.invoke(handle)).reinterpret(512);
            throw new RuntimeException("plife status " + status + ": " + msg.getString(0));
        }
    }
}
