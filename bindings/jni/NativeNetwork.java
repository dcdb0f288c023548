// Reference-side binding a maintainer would add (core/src/main/java/net/consensys/wittgenstein/core/NativeNetwork.java).
// It is the Java face of include/wtg.h: the state of the network lives on the GPU(s); protocols of the accelerated set
// forward their parameters instead of building Java node objects (INTEGRATION.md §1 lists every call site).
// Not compiled in this repository's image (no JDK); the C half (wtg_jni.c) is syntax-checked against a stub jni.h by
// tests/test_abi.py and built by __graft_entry__.build() when a JDK's jni.h is present.
package net.consensys.wittgenstein.core;

public final class NativeNetwork implements AutoCloseable {
  static {
    System.loadLibrary("wtg_jni"); // links libwtg_b200.so
  }

  private long handle;

  /** new Network<>() */
  public NativeNetwork() {
    handle = create();
  }

  /** shard `rank` of `world` of one network spread over several GPUs (wtg_shard_create), one caller thread per shard */
  public NativeNetwork(int rank, int world, int device) {
    handle = shardCreate(rank, world, device);
  }

  private static native long create();

  private static native long shardCreate(int rank, int world, int device);

  private static native void destroy(long h);

  public native void setSeed(long seed); // network.rd.setSeed

  public native void setNodeBuilder(String registryName); // RegistryNodeBuilders.getByName

  public native void setNetworkLatency(String registryName); // RegistryNetworkLatencies.getByName

  public native void setNetworkLatencyMeasured(int[] distribProp, int[] distribVal); // Network.java:665-667

  public native void setMsgDiscardTime(int ms);

  public native void pingPongInit(int nodeCt); // PingPong.init()

  public native void gsfInit(int[] params7); // GSFSignatureParameters fields in declaration order

  public native void handelInit(int[] params11); // HandelParameters

  public native void sanFerminConstruct(int[] params6);

  public native void sanFerminInit();

  public native void capposInit(int[] params6);

  public native void casperConstruct(int[] params6);

  public native void casperInit(int byzantineDelay);

  public native boolean runMs(int ms); // throws IllegalStateException where the reference throws

  public native int time();

  public native int msgsSize();

  public native int msgsSizeAt(int t);

  /**
   * network.msgs.peekMessages() (Network.java:279-286): one row of {from, to, sentAt, arrivingAt, kind, msgType} per
   * pending arrival, sorted by arrival time, flattened (6 ints per row) — the data of EnvelopeInfo (EnvelopeInfo.java:8-14).
   */
  public native int[] peekMessages(int maxRows);

  public native void stopNode(int id);

  public native void startNode(int id);

  public native void partition(float part);

  public native void endPartition();

  /** 5 x count values: msgReceived, msgSent, bytesSent, bytesReceived, doneAt of this engine's nodes */
  public native long[] nodeCounters();

  /** verifiedSignatures of this engine's nodes: count x N/64 words */
  public native long[] gsfVerified();

  public native int[] pingPongPongs();

  /** node-sharded networks: 128-byte handle of this shard's exchange region / mapping of all shards' regions */
  public native byte[] shardExport();

  public native void shardLink(byte[] handlesOfAllShards);

  /** { first id, count } of the nodes this engine owns */
  public native int[] shardRange();

  @Override
  public void close() {
    if (handle != 0) destroy(handle);
    handle = 0;
  }
}
