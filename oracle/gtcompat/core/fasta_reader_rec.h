/* gtcompat: callback-driven FASTA reader (GenomeTools core/fasta_reader.h +
   fasta_reader_rec.h surface used by parser.c:497-549). */
#ifndef GTCOMPAT_FASTA_READER_REC_H
#define GTCOMPAT_FASTA_READER_REC_H
#include "core/error.h"
#include "core/str_api.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GtFastaReader GtFastaReader;
typedef int (*GtFastaReaderProcDescription)(const char *description,
                                            GtUword length, void *data,
                                            GtError *err);
typedef int (*GtFastaReaderProcSequencePart)(const char *seqpart,
                                             GtUword length, void *data,
                                             GtError *err);
typedef int (*GtFastaReaderProcSequenceLength)(GtUword length, void *data,
                                               GtError *err);
GtFastaReader *gt_fasta_reader_rec_new(GtStr *sequence_filename);
int gt_fasta_reader_run(GtFastaReader *reader,
                        GtFastaReaderProcDescription proc_description,
                        GtFastaReaderProcSequencePart proc_sequence_part,
                        GtFastaReaderProcSequenceLength proc_sequence_length,
                        void *data, GtError *err);
void gt_fasta_reader_delete(GtFastaReader *reader);
#ifdef __cplusplus
}
#endif
#endif
