package gammaspike.tree;

import beast.base.evolution.tree.Tree;
import beast.base.util.Randomizer;
import gammaspike.b200.GsmContext;

/** Stubs whose log-time sampling (stub-free inference mode) runs on the device for the whole tree at once (UNCOMPILED here).
 *  Randomizer stays the source of randomness: one nextDouble() per branch that reaches Randomizer.randomChoice in the
 *  reference (non-root, Poisson mean != 0), drawn in node order as Stubs.sampleNStubs does (Stubs:389-401), so a seeded run
 *  consumes the same random numbers in the same order. */
public class StubsB200 extends Stubs {
    private long sampledFor = Long.MIN_VALUE;
    private int[] sampled;

    private int[] sampleAll(long sampleNr) {
        if (sampleNr == sampledFor && sampled != null) return sampled;           // Stubs:420-422: one draw per logged state
        final Tree tree = (Tree) treeInput.get();
        final int n = tree.getNodeCount();
        final double[] u = new double[n];
        for (int i = 0; i < n; i++) {
            final beast.base.evolution.tree.Node node = tree.getNode(i);
            if (node.isRoot() || node.getParent().getHeight() <= node.getHeight()) continue;   // no draw (Stubs:416-434; STP:876-878)
            u[i] = Randomizer.nextDouble();
        }
        sampled = GsmContext.of(tree).stubChoose(u);
        sampledFor = sampleNr;
        return sampled;
    }

    /** Stubs:410-458 */
    @Override
    public int sampleNStubsOnBranch(int nodeNr, long sampleNr) {
        if (this.estimateStubs()) return -1;
        return sampleAll(sampleNr)[nodeNr];
    }

    /** Stubs:389-401 */
    @Override
    public int sampleNStubs(long sampleNr) {
        if (this.estimateStubs()) return -1;
        int total = 0;
        for (int k : sampleAll(sampleNr)) total += k;
        return total;
    }
}
