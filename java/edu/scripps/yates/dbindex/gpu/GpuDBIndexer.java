package edu.scripps.yates.dbindex.gpu;

import java.io.IOException;

import edu.scripps.yates.dbindex.DBIndexer;
import edu.scripps.yates.utilities.fasta.dbindex.DBIndexSearchParams;
import edu.scripps.yates.utilities.fasta.dbindex.DBIndexStoreException;

/**
 * DBIndexer with the digestion moved to the GPU: run() still streams the FASTA and fills the
 * ProteinCache (DBIndexer.java:600-616), but cutSeq only hands the protein to the store; the windows,
 * masses and gates of DBIndexer.java:256-394 are computed by digest_count/emit_kernel when
 * stopAddSeq() calls dbi_build. Uses the reference's own plugin constructor (DBIndexer.java:143).
 * NOT COMPILED HERE (no JDK, see DbiNative).
 */
public class GpuDBIndexer extends DBIndexer {
	public GpuDBIndexer(DBIndexSearchParams sparam, IndexerMode mode) {
		super(sparam, mode, new GpuDBIndexStore(sparam));
	}

	@Override
	protected void cutSeq(final String protAccession, String protSeq) throws IOException {
		try {
			indexStore.addProteinDef(++protNum, protAccession, protSeq); // DBIndexer.java:251
		} catch (final DBIndexStoreException e) {
			throw new IOException(e);
		}
	}
}
